"""The reference's OWN formulation of the contact query -- dense tensor algebra that materialises the [V,V]
distance matrix and the [Q,F,3,3] solid-angle operands (tuch/utils/contact.py:23-147, losses.py:76-93) --
restated with torch ops and run on the same B200, next to the product path.  This is the denominator of
north_star's ">= 10x the reference's per-iteration wall-clock" target: the reference cannot travel to the
GPU box (licence, un-vendored dependencies), so its per-body op chain is restated here as TEST code, timed on
a few bodies (the reference loops over bodies, losses.py:74) and compared with the batched kernels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def reference_chain(v, faces, geomask):
    """One body: exterior flags and masked nearest vertex the way the reference computes them."""
    x = v[None]
    # contact.py:23-47 (squared=True): three K=3 bmm's, diagonals, broadcast adds
    xx, yy, zz = torch.bmm(x, x.transpose(2, 1)), torch.bmm(x, x.transpose(2, 1)), torch.bmm(x, x.transpose(2, 1))
    n = v.shape[0]
    idx = torch.arange(n, device=v.device)
    rx = xx[:, idx, idx].unsqueeze(1).expand_as(zz.transpose(2, 1))
    ry = yy[:, idx, idx].unsqueeze(1).expand_as(zz)
    P = rx.transpose(2, 1) + ry - 2 * zz
    # contact.py:49-147: solid angles of every (query, triangle) pair, summed
    tri = v[faces][None]                                            # [1,F,3,3]
    pts = v[None]
    centered = tri[:, None] - pts[:, :, None, None]                 # [1,Q,F,3,3]
    norms = torch.norm(centered, dim=-1)
    cross = torch.cross(centered[:, :, :, 1], centered[:, :, :, 2], dim=-1)
    num = (centered[:, :, :, 0] * cross).sum(-1)
    del cross
    prod = norms.prod(-1)
    d01 = (centered[:, :, :, 0] * centered[:, :, :, 1]).sum(-1)
    d02 = (centered[:, :, :, 0] * centered[:, :, :, 2]).sum(-1)
    d12 = (centered[:, :, :, 1] * centered[:, :, :, 2]).sum(-1)
    del centered
    den = prod + d01 * norms[:, :, :, 2] + d02 * norms[:, :, :, 1] + d12 * norms[:, :, :, 0]
    del d01, d02, d12, norms
    winding = (2 * torch.atan2(num, den)).sum(-1) / (4 * np.pi)
    exterior = winding.squeeze().le(0.99)
    # losses.py:92-93
    P[:, ~geomask] = float('inf')
    return exterior, torch.argmin(P, dim=1)[0], winding.squeeze()


def test_reference_formulation_on_the_same_gpu(full_assets):
    from tuch_b200 import ops
    from test_contact_gpu import posed_verts, make_topology
    a = full_assets
    n_ref, B = 3, 64
    verts = torch.tensor(posed_verts(a, B, seed=61), device=DEV)
    faces = torch.tensor(a['model']['faces'], device=DEV)
    geomask = torch.tensor(a['geo'], device=DEV) > 0.3
    reference_chain(verts[0], faces, geomask)                       # warm-up (allocator, cuBLAS)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ref = [reference_chain(verts[b], faces, geomask) for b in range(n_ref)]
    e1.record()
    torch.cuda.synchronize()
    ref_ms_per_body = e0.elapsed_time(e1) / n_ref
    topo = make_topology(a, DEV, segments=False, regions=False, exact=False)
    topo.contact_query(verts, use_segments=False)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        out = topo.contact_query(verts, use_segments=False)
    e1.record()
    torch.cuda.synchronize()
    ours_ms_per_body = e0.elapsed_time(e1) / 5 / B
    print('contact query per body on this GPU: reference formulation %.2f ms, tuch_b200 %.4f ms (x%.0f)'
          % (ref_ms_per_body, ours_ms_per_body, ref_ms_per_body / ours_ms_per_body))
    for b in range(n_ref):
        ext, am, w = ref[b]
        safe = (w - 0.99).abs() > 1e-4
        assert torch.equal(out['exterior'][b][safe], ext[safe])
        assert int((out['argmin'][b].long() != am).sum()) <= 3       # fp32 near-ties only
    assert ref_ms_per_body / ours_ms_per_body > 10.0
