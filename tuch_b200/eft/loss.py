"""Host-side mirror of the contact term of tuch/eft/loss.py (EFTLoss.contact_loss :129-181): the third
consumer of the self-contact primitives, with its own reductions -- push and pull are MEANS over the
interior / exterior vertices (no euclthres gate on the pull term, :158-166), the region-to-region term
sums the annotated pair minima, and the total is 100 * (contact + 0.5 * r2r) summed over the bodies.

The reference hands the WHOLE batch to batch_has_self_isec inside its per-body loop (:150), which is
only meaningful for batch size 1 (the EFT fitter's setting); here the segment whitelist is evaluated
per body.
"""
import torch

from .. import ops
from ..smplify.losses import topology_for


class _EftContact(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, topo, pair_active):
        B = verts.shape[0]
        g_verts = torch.zeros(B, topo.V, 3, device=verts.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        q = topo.contact_query(verts, use_segments=len(topo.segment_names) > 0)
        contact, _ = ops.contact_loss(verts, q['argmin'], q['exterior'], 0.0, pull_mode=ops.PULL_ALL,
                                      reduce_mode=ops.REDUCE_MEAN, weight=100.0, g_points=g_verts)
        total = 100.0 * contact
        if pair_active is not None:
            mn, ai, aj = topo.region_min(verts, masked=True, active=pair_active)
            r2r = ops.region_sum(verts, mn, ai, aj, weight=50.0, g_verts=g_verts)
            total = total + 50.0 * r2r
        ctx.save_for_backward(g_verts if g_verts is not None else torch.empty(0, device=verts.device))
        return total.sum()

    @staticmethod
    def backward(ctx, g):
        (gv,) = ctx.saved_tensors
        return (gv * g if gv.numel() else None), None, None


def contact_loss(gt_contact, verts, geomask, face_tensor, cdict, segments=None):
    """sum_b 100 * (mean push + mean pull + 0.5 * sum_pairs min masked squared distance)."""
    topo = topology_for(geomask, face_tensor, verts.shape[1], cdict, segments)
    act = None
    if gt_contact is not None and len(topo.classes) > 0:
        act = (gt_contact.to(verts.device) == 1)
    return _EftContact.apply(verts, topo, act)
