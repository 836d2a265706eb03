"""How many vertices are interior before / after the segment whitelist on the bench's bodies, and in how many
vertex tiles (= warps of the nearest-vertex kernel) they sit."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                                        # noqa: E402
from tuch_b200 import ops                           # noqa: E402


class A:
    gpus, steps, warmup, batch = 1, 10, 3, 256


B = 64
rig = bench.Rig(A)
a = bench.make_assets(B, seed=1000)
d = {k: rig.t(a['inp'][k]) for k in bench.INPUT_KEYS}
s = rig.stack(a, B, num_iters=10)
fit = rig.begin(s, a, d)
tree = ops.cluster_tree(a['model']['faces'], a['model']['v_template'])
vt = torch.as_tensor(tree['vtile']).to(rig.dev).long()
pad = vt < 0
done = 0
for it in (1, 20, 60, 100):
    while done < it:
        fit.step()
        done += 1
    v = fit.vertices
    for name, seg in (('before whitelist', False), ('after whitelist', True)):
        q = fit.topo.contact_query(v, use_segments=seg)
        inter = ~q['exterior']
        t_int = (inter[:, vt.clamp(min=0)] & ~pad).any(-1)
        print('iteration %d, %s: interior vertices %.2f %%, tiles with an interior vertex %.1f %%, in contact (< 2 cm) %.2f %%'
              % (it, name, 100 * inter.float().mean(), 100 * t_int.float().mean(), 100 * (q['min_sq'] < 4e-4).float().mean()))
q = fit.topo.contact_query(fit.vertices, use_segments=True)
inter = ~q['exterior']
for r in (0.02, 0.04, 0.06, 0.08, 0.10, 0.15, 0.20):
    print('interior vertices with an allowed vertex within %.2f m: %.1f %%' % (r, 100 * (q['min_sq'][inter] < r * r).float().mean()))
