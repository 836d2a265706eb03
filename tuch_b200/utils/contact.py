"""Host-side mirror of tuch/utils/contact.py: batch_pairwise_dist (:23-47), solid_angles (:49-109),
winding_numbers (:112-147) -- same names, arguments and return shapes, computed by the sm_100a
kernels in libtuch_b200.so (no CPU path)."""
import torch

from .. import ops


class _PairwiseDist(torch.autograd.Function):
    """The output is NOT saved for backward: the reference's own idiom writes into it in place under no_grad
    (`pred_verts_dists[:, ~geomask] = inf`, losses.py:92, eft/loss.py:155) and then differentiates
    torch.min(...) of it.  The squared form does not need P in backward; the sqrt form recomputes it."""

    @staticmethod
    def forward(ctx, x, y, squared):
        P = ops.pairwise_dist(x, y, squared)
        ctx.save_for_backward(x, y)
        ctx.squared = squared
        return P

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        P = None if ctx.squared else ops.pairwise_dist(x, y, False)
        gx, gy = ops.pairwise_dist_backward(x, y, P, g.contiguous(), ctx.squared)
        return gx, gy, None


def batch_pairwise_dist(x, y, use_cuda=True, squared=True):
    """P[b,i,j] = |x_i|^2 + |y_j|^2 - 2 x_i.y_j  (sqrt if not squared).  `use_cuda` is accepted for
    signature parity; the computation always runs on the tensors' CUDA device."""
    if x.requires_grad or y.requires_grad:
        return _PairwiseDist.apply(x, y, squared)
    return ops.pairwise_dist(x, y, squared)


def solid_angles(points, triangles, thresh=1e-8):
    """[B,Q,F] signed solid angles (Van Oosterom & Strackee).  `thresh` is unused, as in the reference."""
    return ops.solid_angles(points, triangles)


def winding_numbers(points, triangles, thresh=1e-8):
    """[B,Q] generalized winding numbers; the [Q,F] solid-angle matrix is never materialised."""
    return ops.winding_numbers(points, triangles)
