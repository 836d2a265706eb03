"""BASELINE config 2: batch=64 SMPLify-DC, 100 iterations per stage, synthetic DSC contact pairs, 1 GPU, through
the public SMPLifyDC.__call__ (stage 1: camera + shape, stage 2: pose with contact).  Prints wall time and
the loss / interior-vertex trajectory; `graph` as argv[1] replays stage 2 as a CUDA graph."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tuch_b200 import ops, synthetic as syn
from tuch_b200.models.smpl import SMPL
from tuch_b200.smplify.prior import MaxMixturePrior
from tuch_b200.smplify.smplifydc import SMPLifyDC
from tuch_b200.utils.segmentation import BatchBodySegment

B, ITERS = 64, 100
use_graph = len(sys.argv) > 1 and sys.argv[1] == 'graph'
dev = torch.device('cuda:0')
model = syn.make_lattice_body_model(seed=0)
geo = syn.make_geodesics(model['v_template'], model['faces'], cache_dir='/tmp/tuch_b200_cache')
regions, segs, gmm = syn.make_regions(model), syn.make_segments(model), syn.make_gmm()
_smpl_cpu_free = SMPL(model_arrays=model, batch_size=B).to(dev)          # keypoint targets from the product SMPL


def _joints(p, b):
    with torch.no_grad():
        o = _smpl_cpu_free(global_orient=torch.tensor(p[:, :3], device=dev), body_pose=torch.tensor(p[:, 3:], device=dev),
                           betas=torch.tensor(b, device=dev))
    return o.joints.cpu().numpy()


inp = syn.make_smplify_inputs(model, regions, B, seed=2,
                              joints_fn=_joints)
t = lambda x: torch.tensor(np.asarray(x), device=dev)
smpl = SMPL(model_arrays=model, batch_size=B).to(dev)
prior = MaxMixturePrior(gmm=gmm, num_gaussians=8).to(dev)
faces = t(model['faces'])
segments = BatchBodySegment(list(segs.keys()), faces, segment_data=segs)
ign = [syn.JOINT_IDS[n] for n in syn.IGN_JOINTS]
opt = SMPLifyDC(step_size=1e-2, batch_size=B, num_iters=ITERS, focal_length=syn.FOCAL_LENGTH, geodistssmpl=t(geo),
                geothres=0.3, euclthres=0.02, device=dev, smpl=smpl, pose_prior=prior, ign_joints=ign,
                use_cuda_graph=use_graph)
args = (t(inp['init_pose']), t(inp['init_betas']), t(inp['init_cam_t']), t(inp['camera_center']), t(inp['keypoints_2d']))
kw = dict(use_contact=True, contactlist=regions, gt_contact=[t(inp['gt_contact']), None], ignore_idxs=t(inp['ignore_idxs']),
          has_discrete_contact=t(inp['has_discrete_contact']), has_gt_keypoints=None, contact_loss_weight=2000.0,
          contact_loss_return='sum', segments=segments)
opt(*args, **kw)                                     # warm-up (scratch arenas, hierarchy, caches)
torch.cuda.synchronize()
n0 = ops.launch_count()
t0 = time.perf_counter()
out = opt(*args, **kw)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
verts, joints, pose, betas, cam_t, reproj, optiverts = out
topo = ops.Topology(model['faces'], len(model['v_template']), dev)
topo.set_template(model['v_template'])
first = int((~topo.contact_query(optiverts[0].detach(), use_segments=False, want_nearest=False)['exterior']).sum())
last = int((~topo.contact_query(verts, use_segments=False, want_nearest=False)['exterior']).sum())
print('config 2 (B=%d, %d + %d iterations, graph=%s): %.3f s wall, %.2f ms per iteration over both stages, %d library launches'
      % (B, ITERS, ITERS, use_graph, dt, dt * 1e3 / (2 * ITERS), ops.launch_count() - n0))
print('interior vertices: %d at the first stage-2 iteration -> %d at the end; reprojection loss mean %.3f; finite %s'
      % (first, last, float(reproj.mean()), bool(torch.isfinite(pose).all())))
