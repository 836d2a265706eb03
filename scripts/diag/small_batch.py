"""Per-iteration time of the fused stage-2 iteration at a small batch: eager and as a CUDA graph (events), for the
strong-scaling analysis.  Under `ncu --metrics gpu__time_duration.sum` the eager steps give the per-kernel list."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                                        # noqa: E402


class A:
    gpus, steps, warmup, batch = 1, 10, 3, 256


rig = bench.Rig(A)
for B in [int(x) for x in sys.argv[1:]] or [32]:
    a = bench.make_assets(B, seed=1000)
    s = rig.stack(a, B, num_iters=10)
    d = {k: rig.t(a['inp'][k]) for k in bench.INPUT_KEYS}
    fit = rig.begin(s, a, d)
    ms_e, _ = rig.timed(fit.step, 20, warmup=5)
    fit_g = rig.begin(s, a, d).capture()
    ms_g, _ = rig.timed(fit_g.step, 20, warmup=5)      # same iterations as the eager fit: later ones do more work
    cam = bench.stage1_fit(rig, s, d)
    ms_1, _ = rig.timed(cam.step, 50, warmup=5)
    print('B=%d  stage-2 eager %.3f ms  graph %.3f ms  | stage-1 graph %.3f ms' % (B, ms_e, ms_g, ms_1))
