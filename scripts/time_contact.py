"""Dev tool: time the contact query on synthetic SMPL-sized bodies (CUDA events)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tuch_b200 import ops, synthetic as syn

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device('cuda:0')
print(ops.device_info())
m = syn.make_lattice_body_model() if os.environ.get('BODY', 'lattice') == 'lattice' else syn.make_body_model()
V = len(m['v_template'])
topo = ops.Topology(m['faces'], V, dev)
t0 = time.time()
rng = np.random.default_rng(0)
# cheap synthetic "geodesic" mask for timing only: euclidean template distance
vt = torch.tensor(m['v_template'], device=dev)
geo = torch.cdist(vt, vt)
topo.set_geodist(geo, 0.3)
verts = (vt[None] + 0.01 * torch.randn(B, V, 3, device=dev)).contiguous()

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

# posed bodies with self-penetration (the product's fused LBS), tiled to the batch
from tuch_b200.models.smpl import SMPL
nb = min(B, 16)
pose = torch.tensor(syn.fold_arms_pose(nb, seed=7))
betas = torch.tensor(np.random.default_rng(7).normal(0, 0.5, size=(nb, 10)).astype(np.float32))
with torch.no_grad():
    pv = SMPL(model_arrays=m, batch_size=nb).to(dev)(global_orient=pose[:, :3].to(dev), body_pose=pose[:, 3:].to(dev),
                                                       betas=betas.to(dev)).vertices
verts = pv.repeat((B + nb - 1) // nb, 1, 1)[:B].contiguous()
topo.set_template(m['v_template'])
print('clusters', topo.cluster_stats())
ops.kernel_timing(enable=True, reset=True)
topo.set_winding_mode(topo.WINDING_EXACT)
tw = timeit(lambda: topo.contact_query(verts, use_segments=False, want_nearest=False))
w_exact = topo.contact_query(verts, use_segments=False, want_nearest=False)
topo.set_winding_mode(topo.WINDING_FAST)
tf = timeit(lambda: topo.contact_query(verts, use_segments=False, want_nearest=False))
w_fast = topo.contact_query(verts, use_segments=False, want_nearest=False)
err = (w_exact['winding'] - w_fast['winding']).abs()
away = (w_exact['winding'] - 0.99).abs() > 1e-4
print('fast winding %.3f ms (exact %.3f ms): max err %.2e, flag mismatches %d, interior %d, near-threshold %d' % (
    tf, tw, float(err.max()), int((w_exact['exterior'] != w_fast['exterior'])[away].sum()), int((~w_exact['exterior']).sum()),
    int(((w_fast['winding'] - 0.99).abs() < 0.02).sum())))
for k in ('winding_kernel', 'winding_refine_kernel', 'nearest_kernel'):
    print(' ', k, ops.kernel_time(k))
tn = timeit(lambda: topo.contact_query(verts, use_segments=False, want_winding=False))
n_fast = topo.contact_query(verts, use_segments=False, want_winding=False)
topo_old = ops.Topology(m['faces'], V, dev)
topo_old.set_geodist(geo, 0.3)
topo_old.set_winding_mode(topo_old.WINDING_EXACT)
t_old = timeit(lambda: topo_old.contact_query(verts, use_segments=False, want_winding=False))
n_old = topo_old.contact_query(verts, use_segments=False, want_winding=False)
print('nearest: tiles %.3f ms, dense %.3f ms, identical %s' % (tn, t_old, bool(torch.equal(n_fast['argmin'], n_old['argmin']) and torch.equal(n_fast['min_sq'], n_old['min_sq']))))
pairs = B * V * 13776
print('B=%d winding %.3f ms (%.1f Gpairs/s)  nearest %.3f ms (%.1f Gpairs/s)' % (B, tw, pairs / tw / 1e6, tn, B * V * V / tn / 1e6))
