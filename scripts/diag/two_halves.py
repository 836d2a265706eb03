"""Two half-batches fitted concurrently on two streams against one fit of the whole batch: ms per iteration of all
bodies.  The small latency-bound kernels at the head and tail of one half's iteration overlap with the other half's
big kernels.  Usage: python scripts/diag/two_halves.py [B] [offset_ms]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                                        # noqa: E402


class A:
    gpus, steps, warmup, batch = 1, 10, 3, 256


B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
offset_ms = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
rig = bench.Rig(A)
a = bench.make_assets(B, seed=1000)
d = {k: rig.t(a['inp'][k]) for k in bench.INPUT_KEYS}
s = rig.stack(a, B, num_iters=10)
fit = rig.begin(s, a, d)
ms_whole, _ = rig.timed(fit.step, 20, warmup=5)
print('B=%d one fit: %.3f ms per iteration' % (B, ms_whole))

h = B // parts
streams = [torch.cuda.Stream() for _ in range(parts)]
fits = []
for i, st in enumerate(streams):
    di = {k: v[i * h:(i + 1) * h].contiguous() for k, v in d.items()}
    si = rig.stack(a, h, num_iters=10)
    with torch.cuda.stream(st):
        fits.append(rig.begin(si, a, di))
torch.cuda.synchronize()


def step_all():
    for f, st in zip(fits, streams):
        with torch.cuda.stream(st):
            f.step()


def run(steps):
    cur = torch.cuda.current_stream()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for i, st in enumerate(streams):
        st.wait_stream(cur)
        if offset_ms > 0 and i > 0:
            with torch.cuda.stream(st):
                torch.cuda._sleep(int(offset_ms * i * 1.9e6))
    for _ in range(steps):
        step_all()
    for st in streams:
        cur.wait_stream(st)
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / steps


run(5)
print('B=%d as %d concurrent fits of %d (offset %.2f ms): %.3f ms per iteration' % (B, parts, h, offset_ms, run(40) - offset_ms * (parts - 1) / 40))
# the fitted parameters do not depend on the split
ref = rig.begin(s, a, d)
for _ in range(3):
    ref.step()
chk = []
for i, st in enumerate(streams):
    di = {k: v[i * h:(i + 1) * h].contiguous() for k, v in d.items()}
    with torch.cuda.stream(st):
        f = rig.begin(rig.stack(a, h, num_iters=10), a, di)
        for _ in range(3):
            f.step()
        chk.append(f.body_pose.clone())
torch.cuda.synchronize()
print('pose after 3 iterations bit-identical to the whole batch:', bool(torch.equal(torch.cat(chk), ref.body_pose)))
