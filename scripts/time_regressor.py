"""Dev tool: time RegressorLoss.contact_loss's fused path (tuch_regressor_contact_loss, forward + vertex
gradient) on SMPL-sized bodies with a 20k-point HD model."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tuch_b200 import ops, synthetic as syn
from tuch_b200.utils.segmentation import BatchBodySegment
from tuch_b200.models.smpl import SMPL

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device('cuda:0')
m = syn.make_lattice_body_model()
V = len(m['v_template'])
geo = syn.make_geodesics(m['v_template'], m['faces'], cache_dir='/tmp/tuch_b200_cache')
segs = syn.make_segments(m)
hd_reg, hd_fidx = syn.make_hd_regressor(m, n_hd=20000)
topo = ops.Topology(m['faces'], V, dev)
topo.set_template(m['v_template'])
topo.set_geodist(torch.tensor(geo, device=dev), 0.3)
bbs = BatchBodySegment(list(segs.keys()), torch.tensor(m['faces'], device=dev), segment_data=segs)
topo.set_segments(bbs.topology_entries())
topo.set_hd(hd_reg, hd_fidx)
nb = min(B, 16)
pose = torch.tensor(syn.fold_arms_pose(nb, seed=7))
betas = torch.tensor(np.random.default_rng(7).normal(0, 0.5, size=(nb, 10)).astype(np.float32))
with torch.no_grad():
    pv = SMPL(model_arrays=m, batch_size=nb).to(dev)(global_orient=pose[:, :3].to(dev), body_pose=pose[:, 3:].to(dev),
                                                       betas=betas.to(dev)).vertices
verts = pv.repeat((B + nb - 1) // nb, 1, 1)[:B].contiguous()
g = torch.zeros_like(verts)

ops.kernel_timing(enable=True, reset=True)
def run(use_hd):
    return topo.regressor_contact_loss(verts, euclthres=0.02, use_hd=use_hd, g_verts=g)

for use_hd in (False, True):
    run(use_hd); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): loss = run(use_hd)
    e.record(); torch.cuda.synchronize()
    print('B=%d use_hd=%s: %.3f ms per call, mean loss %.4f' % (B, use_hd, s.elapsed_time(e) / 5, float(loss.mean())))

# hierarchical vs all-faces inside test of the HD points: flags and loss
topo.set_winding_mode(topo.WINDING_EXACT)
le, de = topo.regressor_contact_loss(verts[:16].contiguous(), euclthres=0.02, use_hd=True, debug=True)
topo.set_winding_mode(topo.WINDING_FAST)
lf, df = topo.regressor_contact_loss(verts[:16].contiguous(), euclthres=0.02, use_hd=True, debug=True)
n = df['counts']
mism = sum(int((de['hd_exterior'][b, :int(n[b])] != df['hd_exterior'][b, :int(n[b])]).sum()) for b in range(16))
print('selected HD points per body', n.tolist()[:8], 'interior frac %.3f' % float(sum(int((df['hd_exterior'][b, :int(n[b])] == 0).sum()) for b in range(16)) / float(n.sum())))
print('HD exterior flag mismatches fast vs exact: %d of %d; loss rel diff %.2e' % (mism, int(n.sum()), float(((le - lf).abs() / le.abs().clamp_min(1e-9)).max())))
for k in ('winding_kernel_points', 'winding_refine_kernel'):
    print(' ', k, ops.kernel_time(k))

# device time of every timed kernel group over one more HD call
ops.kernel_timing(enable=True, reset=True)
run(True); torch.cuda.synchronize()
for k, v in sorted(ops.kernel_times().items(), key=lambda kv: -kv[1][0]):
    print('  %-28s %8.3f ms  (%d launches)' % (k, v[0], v[1]))
