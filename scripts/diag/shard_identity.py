"""Is a fit of N bodies as one batch bit-identical to the same bodies fitted as shards (what BASELINE config 4 does
across GPUs)?  One GPU; prints the first stage / iteration count at which the results differ."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                                        # noqa: E402


class A:
    gpus, steps, warmup, batch = 1, 10, 3, 256


rig = bench.Rig(A)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
a = bench.make_assets(N, seed=4)
for iters in (1, 3, 10):
    full = {k: rig.t(a['inp'][k]) for k in bench.INPUT_KEYS}
    s1 = rig.stack(a, N, num_iters=iters)
    o1 = s1['smplify'](full['init_pose'], full['init_betas'], full['init_cam_t'], full['camera_center'], full['keypoints_2d'],
                       **rig.call_args(s1, a, full))
    per = N // parts
    outs = []
    for p in range(parts):
        d = {k: v[p * per:(p + 1) * per].contiguous() for k, v in full.items()}
        sp = rig.stack(a, per, num_iters=iters)
        outs.append(sp['smplify'](d['init_pose'], d['init_betas'], d['init_cam_t'], d['camera_center'], d['keypoints_2d'],
                                  **rig.call_args(sp, a, d)))
    names = ['vertices', 'joints', 'pose', 'betas', 'cam_t', 'reproj']
    for i, n in enumerate(names):
        cat = torch.cat([o[i] for o in outs])
        diff = (cat - o1[i]).abs()
        bad = (diff.reshape(N, -1).max(dim=1)[0] > 0).sum().item()
        print('iters %2d %-9s max|diff| %.3e  bodies differing %d / %d' % (iters, n, diff.max().item(), bad, N))
