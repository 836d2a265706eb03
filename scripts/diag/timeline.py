"""Kernel timeline of one fused stage-2 iteration (CUPTI through torch.profiler): name, stream, start and duration
of every kernel, relative to the first kernel of the iteration.  Usage: python scripts/diag/timeline.py [B] [graph]"""
import json
import os
import sys
import tempfile

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                                        # noqa: E402


class A:
    gpus, steps, warmup, batch = 1, 10, 3, 256


B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
graph = len(sys.argv) > 2 and sys.argv[2] == 'graph'
rig = bench.Rig(A)
a = bench.make_assets(B, seed=1000)
d = {k: rig.t(a['inp'][k]) for k in bench.INPUT_KEYS}
s = rig.stack(a, B, num_iters=10)
fit = rig.begin(s, a, d)
if graph:
    fit = fit.capture()
for _ in range(5):
    fit.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        fit.step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), 'trace.json')
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy')]
ev.sort(key=lambda e: e['ts'])
# the second of the three iterations: from its lbs_pose_kernel to the next one
starts = [i for i, e in enumerate(ev) if 'lbs_pose_kernel' in e['name']]
i0, i1 = starts[1], starts[2]
t0 = ev[i0]['ts']
print('B=%d: iteration = %.1f us' % (B, ev[i1]['ts'] - t0))
print('%9s %9s %9s  %-6s %s' % ('start', 'dur', 'end', 'stream', 'kernel'))
for e in ev[i0:i1]:
    name = e['name'].split('(')[0].replace('tuch::', '')
    print('%9.1f %9.1f %9.1f  %-6s %s' % (e['ts'] - t0, e['dur'], e['ts'] - t0 + e['dur'], e['args'].get('stream', '?'), name[:60]))
