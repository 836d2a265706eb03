"""TEST INFRASTRUCTURE -- restatement of tuch/eft/loss.py:129-181 (EFTLoss.contact_loss): the third consumer
of the self-contact primitives.  Pinned by tests/golden/eft_contact_loss.npz, recorded from the reference's own
tuch/eft/loss.py by tests/golden/make_golden_eft.py (tests/test_oracle_golden.py::test_eft_contact_loss_matches_reference)."""
import numpy as np
import torch

from . import losses as ol


def eft_contact_loss(gt_contact, verts, faces, geomask_np, cdict, segments):
    """sum_b 100 * (mean push + mean pull + 0.5 * sum_pairs min masked squared distance)   (eft/loss.py:178-180).
    verts[B,V,3] torch (autograd); gt_contact[B,n_classes]; the segment whitelist always runs (:150-152) and is
    evaluated per body (the reference's whole-batch call at :150 is only defined for B = 1)."""
    total = verts.new_zeros(())
    for b in range(verts.shape[0]):
        vb = verts[b]
        ext, am, _, _ = ol.contact_query(vb, faces, geomask_np, segments, always_segments=True)    # :143-156
        d = torch.norm(vb - vb[torch.as_tensor(am, dtype=torch.long)], dim=1)                       # :159
        e = torch.as_tensor(ext)
        pull = (0.005 * torch.tanh(d[e] / 0.005) ** 2).mean() if e.any() else d.new_zeros(())       # :162-163
        push = (torch.tanh(d[~e] / 0.04) ** 2).mean() if (~e).any() else d.new_zeros(())            # :164-165
        active = np.where(np.asarray(gt_contact[b]) == 1)[0]                                        # :170-171
        r2r = ol.r2r_term(vb, geomask_np, cdict, active) if len(active) else 0.0                    # :172-177
        total = total + 100 * (push + pull + 0.5 * r2r)                                             # :179
    return total
