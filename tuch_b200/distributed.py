"""One process per GPU: sharding of the body batch and the few collectives the path needs.

SMPLify-DC (tuch/smplify/smplifydc.py) optimises every body independently -- the loss is a plain sum
over bodies (losses.py:123) and Adam is element-wise -- so the batch shards across ranks with NO
collective inside the optimisation; results are gathered once at the end.  The train step needs a
gradient all-reduce for the regressor plus a count correction, because RegressorLoss averages over
data-dependent subsets of the GLOBAL batch (`contact_loss[valid_fit].mean()`, loss.py:317; the SPIN
terms :182,:210-236): averaging per-rank means is only right when every rank holds the same number of
valid bodies.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank=None, world_size=None):
    """[lo, hi) of this rank's contiguous shard of n bodies; sizes differ by at most one."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, extra = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(t, rank=None, world_size=None):
    """This rank's slice of a batch-leading tensor (or of every tensor in a list / tuple / dict)."""
    if isinstance(t, dict):
        return {k: shard(v, rank, world_size) for k, v in t.items()}
    if isinstance(t, (list, tuple)):
        return type(t)(shard(v, rank, world_size) for v in t)
    if t is None or not hasattr(t, 'shape') or len(t.shape) == 0:
        return t
    lo, hi = shard_bounds(t.shape[0], rank, world_size)
    return t[lo:hi]


def gather_bodies(t, n_total):
    """Concatenates the per-rank shards of a batch-leading tensor in rank order on every rank."""
    rank, ws = world()
    if ws == 1:
        return t
    sizes = [shard_bounds(n_total, r, ws) for r in range(ws)]
    biggest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((biggest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


def global_masked_mean(per_body, mask):
    """mean of per_body over mask across ALL ranks (differentiable w.r.t. per_body): the single-process
    value of `x[mask].mean()` on the concatenated batch.  Each rank returns local_sum / global_count, so
    that summing the ranks' gradients (what the gradient all-reduce of DDP does with op=SUM) reproduces the
    single-GPU gradient."""
    m = mask.to(per_body.dtype)
    count = m.sum().detach().clone()
    total = (per_body * m).sum()
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(count, op=dist.ReduceOp.SUM)
    return total / count, count


def all_reduce_sum_(tensors):
    """In-place SUM all-reduce of a list of gradient tensors as ONE flattened bucket."""
    rank, ws = world()
    tensors = [t for t in tensors if t is not None]
    if ws == 1 or not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    o = 0
    for t in tensors:
        t.copy_(flat[o:o + t.numel()].view_as(t))
        o += t.numel()
