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


def _spin_terms(crit, d, lo, hi):
    """keypoint / 3-D keypoint / shape / pose / betas / camera terms of RegressorLoss on rows [lo, hi)."""
    sl = lambda k: d[k][lo:hi]
    kp = crit.keypoint_loss(sl('pred_kp'), sl('gt_kp'), 0.5, 1.0, sl('valid'))
    kp3 = crit.keypoint_3d_loss(sl('pred_j'), sl('gt_j'), sl('has3d'))
    sh = crit.shape_loss(sl('pred_v'), sl('opt_v'), sl('valid'))
    po, be = crit.smpl_losses(sl('pred_rot'), sl('pred_betas'), sl('gt_pose'), sl('gt_betas'), sl('valid'), sl('valid'))
    return [kp.reshape(()), kp3.reshape(()), sh.reshape(()), po.reshape(()), be.reshape(())]


def _loss_worker(rank, ws, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    import torch.nn as nn
    from tuch_b200 import distributed as D
    from tuch_b200.train.loss import RegressorLoss, masked_mean
    n = 7
    g = torch.Generator().manual_seed(1)
    r = lambda *s: torch.randn(*s, generator=g)
    d = dict(pred_kp=r(n, 49, 2), gt_kp=torch.cat([r(n, 49, 2), torch.rand(n, 49, 1, generator=g)], -1),
             pred_j=r(n, 49, 3), gt_j=torch.cat([r(n, 24, 3), torch.rand(n, 24, 1, generator=g)], -1),
             pred_v=r(n, 30, 3), opt_v=r(n, 30, 3), pred_rot=r(n, 24, 3, 3), pred_betas=r(n, 10),
             gt_pose=0.3 * r(n, 72), gt_betas=r(n, 10),
             valid=torch.tensor([True, False, True, True, False, False, True]),     # 3 valid on rank 0, 1 on rank 1
             has3d=torch.tensor([0, 0, 0, 0, 1, 1, 0]))                             # none on rank 0
    leaves = ('pred_kp', 'pred_j', 'pred_v', 'pred_rot', 'pred_betas')

    def make(mode):
        crit = RegressorLoss.__new__(RegressorLoss)          # the SPIN terms are plain torch: no CUDA topology needed
        nn.Module.__init__(crit)
        crit.device = torch.device('cpu')
        crit.criterion_shape, crit.criterion_regr = nn.L1Loss(), nn.MSELoss()
        crit.criterion_keypoints = nn.MSELoss(reduction='none')
        crit.dist_mode = mode
        return crit

    # single-process reference on the whole batch
    ref_d = {k: (v.clone().requires_grad_(True) if k in leaves else v) for k, v in d.items()}
    ref_terms = _spin_terms(make(None), ref_d, 0, n)
    torch.stack(ref_terms).sum().backward()
    lo, hi = D.shard_bounds(n)
    for mode in ('sum', 'mean'):
        my_d = {k: (v.clone().requires_grad_(True) if k in leaves else v) for k, v in d.items()}
        terms = torch.stack(_spin_terms(make(mode), my_d, lo, hi))
        terms.sum().backward()
        tot = terms.detach().clone()
        dist.all_reduce(tot)
        if mode == 'mean':
            tot /= ws
        assert torch.allclose(tot, torch.stack(ref_terms).detach(), rtol=1e-5, atol=1e-7), (mode, tot, ref_terms)
        for k in leaves:
            gfull = torch.zeros_like(d[k])
            gfull[lo:hi] = my_d[k].grad[lo:hi]
            dist.all_reduce(gfull)
            if mode == 'mean':
                gfull /= ws
            assert torch.allclose(gfull, ref_d[k].grad, rtol=1e-5, atol=1e-8), (mode, k)
    # a subset that is empty on EVERY rank gives 0 like the reference's `if len == 0` branch
    none = torch.zeros(n, dtype=torch.bool)
    z = make('sum').shape_loss(d['pred_v'][lo:hi], d['opt_v'][lo:hi], none[lo:hi])
    assert float(z) == 0.0
    assert torch.isnan(masked_mean(torch.ones(3), torch.zeros(3, dtype=torch.bool)))        # reference semantics kept
    if rank == 0:
        out.put('ok')
    dist.destroy_process_group()


def test_regressor_loss_terms_count_corrected_world2():
    """RegressorLoss.set_distributed(): every rank's share of the SPIN terms adds up to the single-process value
    and the reduced gradients equal the single-process gradients, with unequal subset sizes per rank."""
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_loss_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
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
