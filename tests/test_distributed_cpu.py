"""Host-side multi-rank logic on CPU (gloo, world_size 2): sharding, gathering in rank order and the
count-corrected global mean of the regressor loss.  No GPU compute is involved."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, n, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    from tuch_b200 import distributed as D
    torch.manual_seed(0)
    full = torch.randn(n, 5)
    mask = torch.tensor([True, False, True, True, False, True, True][:n])
    mine = D.shard(full)
    lo, hi = D.shard_bounds(n)
    assert torch.equal(mine, full[lo:hi])
    d = D.shard({'a': full, 'b': [full[:, 0], None], 's': torch.tensor(3.0)})
    assert torch.equal(d['b'][0], full[lo:hi, 0]) and d['b'][1] is None and float(d['s']) == 3.0
    back = D.gather_bodies(mine * 2, n)
    assert torch.equal(back, full * 2)
    # count-corrected mean: identical to the single-process masked mean, and so is the summed gradient
    x = full[lo:hi, 0].clone().requires_grad_(True)
    val, count = D.global_masked_mean(x ** 2, mask[lo:hi])
    val.backward()
    tot = val.detach().clone()
    dist.all_reduce(tot)
    g = torch.zeros(n)
    g[lo:hi] = x.grad
    D.all_reduce_sum_([g])
    ref_x = full[:, 0].clone().requires_grad_(True)
    ref = (ref_x ** 2)[mask].mean()
    ref.backward()
    assert int(count) == int(mask.sum())
    assert torch.allclose(tot, ref.detach(), atol=1e-6)
    assert torch.allclose(g, ref_x.grad, atol=1e-6)
    if rank == 0:
        out.put('ok')
    dist.destroy_process_group()


def test_shard_gather_and_global_mean_world2():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 'ok'


def test_shard_bounds_cover_batch():
    from tuch_b200 import distributed as D
    for n in (0, 1, 7, 256, 1000):
        for ws in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
