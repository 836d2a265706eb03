"""BASELINE configs 3 and 5 (shape): TUCH train steps on SMPL-sized synthetic batches through the train-step
mirror (tuch_b200.train.train_module.TUCH.forward_train_step + backward + optimiser step), one process per GPU.
The image regressor is a stand-in (HMR / ResNet-50 is outside this path and is plain torch in the reference);
what is timed is everything between its output and its gradient: SMPL forwards, region contact, camera
estimates, SMPLify-DC in the loop (config 5), the fits store, RegressorLoss with the HD contact term, backward.
With WORLD_SIZE > 1 the regressor gradients are all-reduced over NCCL like DistributedDataParallel would.
    python scripts/run_config5.py [bodies_per_gpu=128] [smplify_iters=10] [steps=3]     # smplify_iters=0 -> config 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/run_config5.py 128 10
"""
import os
import sys
from collections import namedtuple

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from tuch_b200 import distributed as tdist, ops, synthetic as syn
from tuch_b200.models.smpl import SMPL
from tuch_b200.smplify.prior import MaxMixturePrior
from tuch_b200.smplify.smplifydc import SMPLifyDC
from tuch_b200.train.fits_dict import FitsDict
from tuch_b200.train.loss import RegressorLoss
from tuch_b200.train.train_module import TUCH
from tuch_b200.utils.segmentation import BatchBodySegment

PER_GPU = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 10
STEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 3
PROFILE = len(sys.argv) > 4 and sys.argv[4] == 'profile'       # synchronised per-phase wall times (dev aid)
SEED = int(os.environ.get('TUCH_BATCH_SEED', 5))
GRAPH = os.environ.get('TUCH_FIT_GRAPH', '1') != '0'          # fitting iterations as CUDA graphs (captured once)
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)

Opt = namedtuple('Opt', ['batch_size', 'img_res', 'run_smplify', 'use_contact_in_the_loop',
                         'contact_in_the_loop_loss_weight', 'smplify_threshold', 'contact_loss_weight',
                         'openpose_train_weight', 'gt_train_weight', 'shape_loss_weight', 'keypoint_loss_weight',
                         'pose_loss_weight', 'beta_loss_weight'])
# train_options.py defaults, contact_loss_weight 1.0 (the 1e-5 default makes the term numerically invisible)
o = Opt(PER_GPU, 224, ITERS > 0, True, 2000.0, 100.0, 1.0, 0.0, 1.0, 0.0, 5.0, 1.0, 0.001)

model = syn.make_lattice_body_model(seed=0)
V = len(model['v_template'])
geo = syn.make_geodesics(model['v_template'], model['faces'], cache_dir='/tmp/tuch_b200_cache')
regions, segs, gmm = syn.make_regions(model), syn.make_segments(model), syn.make_gmm()
hd_reg, hd_fidx = syn.make_hd_regressor(model, n_hd=20000)
t = lambda x: torch.tensor(np.asarray(x), device=dev)
faces = t(model['faces'])
geod = t(geo)
segments = BatchBodySegment(list(segs.keys()), faces, segment_data=segs)
smpl = SMPL(model_arrays=model, batch_size=PER_GPU).to(dev)
face_tensor = faces[None].expand(PER_GPU, -1, -1)
crit = RegressorLoss(o, dev, V, face_tensor, geod, geothres=0.3, euclthres=0.02, face_tensor=face_tensor, use_hd=True,
                     hd_regressor=hd_reg, hd_faces=hd_fidx, segments=segments, template=model['v_template'])
smplify = SMPLifyDC(step_size=1e-2, batch_size=PER_GPU, num_iters=max(ITERS, 1), focal_length=syn.FOCAL_LENGTH,
                    geodistssmpl=geod, geothres=0.3, euclthres=0.02, device=dev, use_cuda_graph=GRAPH,
                    smpl=SMPL(model_arrays=model, batch_size=PER_GPU).to(dev),
                    pose_prior=MaxMixturePrior(gmm=gmm, num_gaussians=8).to(dev),
                    ign_joints=[syn.JOINT_IDS[n] for n in syn.IGN_JOINTS])


def _joints(p, b):
    with torch.no_grad():
        return smpl(global_orient=t(p[:, :3]), body_pose=t(p[:, 3:]), betas=t(b)).joints.cpu().numpy()


batch, store = syn.make_train_batch(model, regions, PER_GPU, seed=SEED + rank, joints_fn=_joints, img_hw=64)
fits = FitsDict(device=dev, dataset_sizes={'dsc': len(store)})
fits.fits_dict['dsc'] = torch.tensor(store)
net = syn.make_stand_in_regressor(seed=0).to(dev)
optim = torch.optim.Adam(net.parameters(), lr=1e-5)
tuch = TUCH(o, dev, None, smpl, None, net, smplify, crit, geod, fits_dict=fits, contactlists=regions,
            focal_length=syn.FOCAL_LENGTH, geothres=0.3, euclthres=0.02)
gb = {k: (v if k == 'dataset_name' else t(v)) for k, v in batch.items()}


def step():
    loss, losses, out = tuch.forward_train_step(gb)
    optim.zero_grad()
    loss.backward()
    if world > 1:                                      # one flattened bucket over NCCL, then the mean
        grads = [p.grad for p in net.parameters()]
        tdist.all_reduce_sum_(grads)
        for g in grads:
            g /= world
    optim.step()
    return losses, out


if PROFILE:
    import time
    phases = {}

    def timed(name, fn):
        def wrapper(*a, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn(*a, **k)
            torch.cuda.synchronize()
            phases[name] = phases.get(name, 0.0) + time.perf_counter() - t0
            return r
        return wrapper
    tuch.smplify = type('Fit', (), {'__call__': staticmethod(timed('smplify.__call__', smplify.__call__)),
                                    'get_fitting_loss': staticmethod(timed('get_fitting_loss', smplify.get_fitting_loss))})()
    tuch.criterion_cospin = type('Crit', (), {'__call__': staticmethod(timed('criterion', crit.__call__)),
                                              'segments': crit.segments})()
    tuch.smpl = timed('smpl forwards', smpl)
    tuch.model = type('Net', (), {'__call__': staticmethod(timed('regressor', net.__call__)), 'train': net.train})()
    fits.__class__ = type('TimedFits', (FitsDict,), {'__getitem__': timed('fits get', FitsDict.__getitem__),
                                                     '__setitem__': timed('fits set', FitsDict.__setitem__)})
    tuch.contact_from_verts = timed('contact_from_verts', tuch.contact_from_verts)
    import tuch_b200.train.train_module as tmod
    tmod.estimate_translation = timed('estimate_translation', tmod.estimate_translation)
    tmod.rotation_matrix_to_angle_axis = timed('rotmat_to_angle_axis', tmod.rotation_matrix_to_angle_axis)
    _step, _bwd = tuch.forward_train_step, torch.Tensor.backward
    tuch.forward_train_step = timed('forward_train_step (all)', _step)
    torch.Tensor.backward = timed('backward', _bwd)

step()                                                 # warm-up: scratch arenas, hierarchy, cuBLAS handles
torch.cuda.synchronize()
if PROFILE:
    phases.clear()
if world > 1:
    dist.barrier()
n0 = ops.launch_count()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(STEPS):
    losses, out = step()
e.record()
torch.cuda.synchronize()
ms = torch.tensor([s.elapsed_time(e) / STEPS], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(ms)
    print('config %s: %d bodies/GPU x %d GPU(s), %d SMPLify-DC iterations per stage in the loop: %.1f ms per train step '
          '(max over ranks) = %.0f bodies/s; %d host-side launches of our kernels per step (graph replays excluded)'
          % ('5' if ITERS > 0 else '3', PER_GPU, world, ITERS, ms, PER_GPU * world / ms * 1e3,
             (ops.launch_count() - n0) // STEPS))
    print('  losses:', {k: round(float(v), 5) for k, v in losses.items()})
    print('  valid fits %d / %d, fits-store rows rewritten so far: %d' % (
        int(out['valid_kpts_anno'].sum()), PER_GPU, int((fits.fits_dict['dsc'] != torch.tensor(store)).any(dim=1).sum())))
    if PROFILE:
        print('  phases, ms per step (synchronised, so they add up to more than the pipelined step):')
        for k, v in sorted(phases.items(), key=lambda kv: -kv[1]):
            print('    %-28s %.2f' % (k, v * 1e3 / STEPS))
if world > 1:
    dist.destroy_process_group()
