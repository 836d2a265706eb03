"""BASELINE config 4 (shape): SMPLify-DC sharded over the GPUs of one box with NCCL -- every rank fits its
contiguous shard of the batch through SMPLifyDC.__call__ (no collective inside the optimisation), the results
are gathered once with all_gather (tuch_b200.distributed.gather_bodies), and rank 0 checks them against the
same batch fitted on its own GPU alone.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/run_config4.py [bodies_per_gpu] [iters]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from tuch_b200 import distributed as tdist, synthetic as syn
from tuch_b200.models.smpl import SMPL
from tuch_b200.smplify.prior import MaxMixturePrior
from tuch_b200.smplify.smplifydc import SMPLifyDC
from tuch_b200.utils.segmentation import BatchBodySegment

PER_GPU = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
N = PER_GPU * world
model = syn.make_lattice_body_model(seed=0)
geo = syn.make_geodesics(model['v_template'], model['faces'], cache_dir='/tmp/tuch_b200_cache')
regions, segs, gmm = syn.make_regions(model), syn.make_segments(model), syn.make_gmm()
_smpl_all = SMPL(model_arrays=model, batch_size=N).to(dev)                     # keypoint targets from the product SMPL


def _joints(p, b):
    with torch.no_grad():
        o = _smpl_all(global_orient=torch.tensor(p[:, :3], device=dev), body_pose=torch.tensor(p[:, 3:], device=dev),
                      betas=torch.tensor(b, device=dev))
    return o.joints.cpu().numpy()


inp = syn.make_smplify_inputs(model, regions, N, seed=4,
                              joints_fn=_joints)
t = lambda x: torch.tensor(np.asarray(x), device=dev)
faces = t(model['faces'])
segments = BatchBodySegment(list(segs.keys()), faces, segment_data=segs)
ign = [syn.JOINT_IDS[n] for n in syn.IGN_JOINTS]
geod = t(geo)


def fit(batch):
    B = len(batch['init_pose'])
    opt = SMPLifyDC(step_size=1e-2, batch_size=B, num_iters=ITERS, focal_length=syn.FOCAL_LENGTH, geodistssmpl=geod,
                    geothres=0.3, euclthres=0.02, device=dev, smpl=SMPL(model_arrays=model, batch_size=B).to(dev),
                    pose_prior=MaxMixturePrior(gmm=gmm, num_gaussians=8).to(dev), ign_joints=ign)
    return opt(t(batch['init_pose']), t(batch['init_betas']), t(batch['init_cam_t']), t(batch['camera_center']),
               t(batch['keypoints_2d']), use_contact=True, contactlist=regions,
               gt_contact=[t(batch['gt_contact']), None], ignore_idxs=t(batch['ignore_idxs']),
               has_discrete_contact=t(batch['has_discrete_contact']), has_gt_keypoints=None,
               contact_loss_weight=2000.0, contact_loss_return='sum', segments=segments)


keys = ('init_pose', 'init_betas', 'init_cam_t', 'camera_center', 'keypoints_2d', 'gt_contact', 'ignore_idxs', 'has_discrete_contact')
mine = tdist.shard({k: inp[k] for k in keys}, rank, world)
fit(mine)                                             # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
verts, joints, pose, betas, cam_t, reproj, _ = fit(mine)
pose_all = tdist.gather_bodies(pose, N)
betas_all = tdist.gather_bodies(betas, N)
verts_all = tdist.gather_bodies(verts, N)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if rank == 0:
    print('config 4: %d bodies over %d GPU(s), %d + %d iterations, fit + all_gather of poses/betas/vertices: %.3f s'
          % (N, world, ITERS, ITERS, dt))
    if world > 1:
        ref = fit({k: inp[k] for k in keys})           # the whole batch on one GPU
        ref2 = fit({k: inp[k] for k in keys})          # ... and once more: the run-to-run noise floor
        d = (ref[2] - pose_all).abs().amax(dim=1)
        print('  sharded vs single-GPU: max |pose diff| %.2e (median over bodies %.2e), max |vertex diff| %.2e'
              % (float(d.max()), float(d.median()), float((ref[0] - verts_all).abs().max())))
        d2 = (ref[2] - ref2[2]).abs().amax(dim=1)
        print('  single-GPU run vs itself:  max |pose diff| %.2e (median %.2e) -- every reduction has a fixed order and '
              'the gradient scatters accumulate in fixed point' % (float(d2.max()), float(d2.median())))
if world > 1:
    dist.destroy_process_group()
