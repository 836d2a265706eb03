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

tw = timeit(lambda: topo.contact_query(verts, use_segments=False, want_nearest=False))
tn = timeit(lambda: topo.contact_query(verts, use_segments=False, want_winding=False))
pairs = B * V * 13776
print('B=%d winding %.3f ms (%.1f Gpairs/s)  nearest %.3f ms (%.1f Gpairs/s)' % (B, tw, pairs / tw / 1e6, tn, B * V * V / tn / 1e6))
