"""GPU parity tests of the contact kernels (through the C ABI) against the CPU oracle and the
golden vectors recorded from the reference."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
# hierarchical winding numbers (tuch_b200/csrc/clusters.h): values within the re-evaluation margin of the 0.99
# threshold are exact; elsewhere the far-field error must stay below a quarter of that margin
WC_MARGIN = 0.10
FAR_FIELD_TOL = 0.25 * WC_MARGIN


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from tuch_b200 import ops
    info = ops.device_info()
    assert info['cc'][0] == 10
    return torch.device('cuda:0')


def posed_verts(assets, batch, seed, dtype=np.float32):
    from oracle import lbs as olbs
    from tuch_b200 import synthetic as syn
    tm = olbs.to_torch_model(assets['model'])
    pose = torch.tensor(syn.fold_arms_pose(batch, seed=seed))
    betas = torch.tensor(np.random.default_rng(seed).normal(0, 0.5, size=(batch, 10)).astype(np.float32))
    return olbs.smpl_forward(tm, betas, pose[:, 3:], pose[:, :3])[0].numpy().astype(dtype)


def make_topology(assets, dev, geothres=0.3, segments=True, regions=True, exact=True, template=True):
    """exact=True pins the winding numbers to the all-faces sum (value parity); exact=False is the
    product default (hierarchical far field + exact re-evaluation near the 0.99 threshold)."""
    from tuch_b200 import ops
    from tuch_b200.utils.segmentation import BatchBodySegment
    m = assets['model']
    topo = ops.Topology(m['faces'], len(m['v_template']), dev)
    topo.set_winding_mode(topo.WINDING_EXACT if exact else topo.WINDING_FAST)
    if not exact and template:
        topo.set_template(m['v_template'])
    topo.set_geodist(torch.tensor(assets['geo'], device=dev), geothres)
    if regions:
        topo.set_regions(assets['regions'])
    if segments:
        bbs = BatchBodySegment(list(assets['segs'].keys()), torch.tensor(m['faces'], device=dev),
                               segment_data=assets['segs'])
        topo.set_segments(bbs.topology_entries())
    return topo


def test_primitives_match_reference_golden(dev):
    from tuch_b200.utils import contact
    g = golden('primitives.npz')
    pts, tri = torch.tensor(g['pts'], device=dev), torch.tensor(g['tri'], device=dev)
    sa = contact.solid_angles(pts, tri).cpu().numpy()
    assert np.abs(sa - g['solid_angles']).max() < 2e-5
    wn = contact.winding_numbers(pts, tri).cpu().numpy()
    assert np.abs(wn - g['winding']).max() < 2e-5
    x, y = torch.tensor(g['x'], device=dev), torch.tensor(g['y'], device=dev)
    assert np.abs(contact.batch_pairwise_dist(x, y).cpu().numpy() - g['pdist_sq']).max() < 2e-5
    pd = contact.batch_pairwise_dist(x, x, squared=False).cpu().numpy()
    ok = np.isfinite(g['pdist']) & (g['pdist'] > 1e-2)
    assert np.abs(pd - g['pdist'])[ok].max() < 1e-4


def test_pairwise_dist_backward_matches_autograd(dev):
    from tuch_b200.utils import contact
    rng = np.random.default_rng(5)
    x = torch.tensor(rng.normal(size=(2, 33, 3)).astype(np.float32), device=dev, requires_grad=True)
    y = torch.tensor(rng.normal(size=(2, 47, 3)).astype(np.float32), device=dev, requires_grad=True)
    w = torch.tensor(rng.normal(size=(2, 33, 47)).astype(np.float32), device=dev)
    for squared in (True, False):
        x.grad = y.grad = None
        (contact.batch_pairwise_dist(x, y, squared=squared) * w).sum().backward()
        gx, gy = x.grad.cpu().double(), y.grad.cpu().double()
        xd, yd = x.detach().cpu().double().requires_grad_(True), y.detach().cpu().double().requires_grad_(True)
        P = (xd * xd).sum(-1)[:, :, None] + (yd * yd).sum(-1)[:, None, :] - 2 * xd @ yd.transpose(1, 2)
        if not squared:
            P = P.sqrt()
        (P * w.cpu().double()).sum().backward()
        assert (gx - xd.grad).abs().max() < 2e-4 * xd.grad.abs().max()
        assert (gy - yd.grad).abs().max() < 2e-4 * yd.grad.abs().max()


def test_pairwise_dist_inplace_mask_idiom(dev):
    """The reference's idiom (losses.py:76,92-93,115; eft/loss.py:155): build P with autograd, write inf into it in
    place under no_grad, then differentiate a min of it.  The mirror must allow that and give torch's gradient."""
    from tuch_b200.utils import contact
    rng = np.random.default_rng(11)
    xn = rng.normal(size=(1, 40, 3)).astype(np.float32)
    mask = rng.uniform(size=(40, 40)) > 0.4
    mask = mask & mask.T
    np.fill_diagonal(mask, False)
    x = torch.tensor(xn, device=dev, requires_grad=True)
    P = contact.batch_pairwise_dist(x, x, squared=True)
    with torch.no_grad():
        P[:, ~torch.tensor(mask, device=dev)] = float('inf')
    sub = P[:, [1, 5, 9], :][:, :, [2, 3, 30, 31]]
    torch.min(sub).backward()
    xd = torch.tensor(xn, dtype=torch.float64, requires_grad=True)
    Pd = (xd * xd).sum(-1)[:, :, None] + (xd * xd).sum(-1)[:, None, :] - 2 * xd @ xd.transpose(1, 2)
    with torch.no_grad():
        Pd[:, ~torch.tensor(mask)] = float('inf')
    torch.min(Pd[:, [1, 5, 9], :][:, :, [2, 3, 30, 31]]).backward()
    assert torch.isfinite(x.grad).all()
    assert (x.grad.cpu().double() - xd.grad).abs().max() < 2e-5 * xd.grad.abs().max()


def test_winding_edge_cases(dev):
    from tuch_b200.utils import contact
    # empty triangle set -> zeros; empty query set -> empty
    p = torch.zeros(2, 5, 3, device=dev)
    assert contact.winding_numbers(p, torch.zeros(2, 0, 3, 3, device=dev)).abs().max() == 0
    assert contact.winding_numbers(torch.zeros(2, 0, 3, device=dev), torch.zeros(2, 4, 3, 3, device=dev)).shape == (2, 0)
    # a query sitting exactly on a triangle corner contributes exactly 0 (atan2(+-0, +0))
    tri = torch.tensor([[[[0., 0, 0], [1, 0, 0], [0, 1, 0]]]], device=dev)
    q = torch.tensor([[[0., 0, 0], [1, 0, 0], [0, 1, 0]]], device=dev)
    assert contact.winding_numbers(q, tri).abs().max() == 0
    # ragged sizes that do not divide any tile
    rng = np.random.default_rng(1)
    from oracle import clib
    for Q, F in ((1, 1), (257, 129), (300, 1000)):
        pts = rng.normal(0, 0.5, size=(Q, 3)).astype(np.float32)
        tr = rng.normal(0, 0.5, size=(F, 3, 3)).astype(np.float32)
        got = contact.winding_numbers(torch.tensor(pts, device=dev)[None], torch.tensor(tr, device=dev)[None])[0].cpu().numpy()
        ref = clib.winding_numbers(pts, tr, dtype=np.float64)
        assert np.abs(got - ref).max() < 3e-5, (Q, F)


def check_query(assets, dev, batch, seed, use_segments):
    from oracle import clib, segments as oseg
    topo = make_topology(assets, dev)
    verts = posed_verts(assets, batch, seed)
    out = topo.contact_query(torch.tensor(verts, device=dev), use_segments=use_segments)
    fast = make_topology(assets, dev, exact=False).contact_query(torch.tensor(verts, device=dev), use_segments=use_segments)
    faces = assets['model']['faces']
    geomask = assets['geo'] > 0.3
    segs = oseg.build_segments(assets['segs'], faces)
    n_interior = 0
    # hierarchical mode: same nearest vertices, winding numbers within the far-field error, flags
    # identical away from the threshold
    assert torch.equal(fast['argmin'], out['argmin']) and torch.equal(fast['min_sq'], out['min_sq'])
    assert (fast['winding'] - out['winding']).abs().max() < FAR_FIELD_TOL
    away = (out['winding'] - 0.99).abs() > 1e-4
    assert torch.equal(fast['exterior'][away], out['exterior'][away])
    for b in range(batch):
        v = verts[b]
        w64 = clib.winding_numbers(v, v[faces], dtype=np.float64)
        w32 = clib.winding_numbers(v, v[faces], dtype=np.float32)
        got_w = out['winding'][b].cpu().numpy()
        assert np.abs(got_w - w64).max() < 2e-5
        assert np.abs(got_w - w32).max() < 2e-5
        ext = w32 <= 0.99
        safe = np.abs(w64 - 0.99) > 1e-4                 # flags must agree away from the threshold
        if use_segments:
            for s in segs:
                sw = clib.winding_numbers(v[s.vidx], s.closed_triangles(v), dtype=np.float64)
                ext[s.vidx[sw > 0.99]] = True
                safe[s.vidx[np.abs(sw - 0.99) <= 1e-4]] = False
        got_e = out['exterior'][b].cpu().numpy()
        assert np.array_equal(got_e[safe], ext[safe])
        n_interior += int((~ext).sum())
        # nearest geodesically-far vertex: identical index, or an fp32 near-tie of the squared distance
        am32, mn32 = clib.masked_nearest(v, geomask, dtype=np.float32)
        got_am = out['argmin'][b].cpu().numpy()
        got_mn = out['min_sq'][b].cpu().numpy()
        diff = np.where(got_am != am32)[0]
        if len(diff):
            P64 = clib.pairwise_dist(v, v, dtype=np.float64)
            for c in diff:
                assert geomask[got_am[c], c]
                assert abs(P64[got_am[c], c] - P64[am32[c], c]) < 2e-6, c
        assert len(diff) <= max(2, batch * len(v) // 500)
        fin = np.isfinite(mn32)
        assert np.array_equal(np.isfinite(got_mn), fin)
        assert np.abs(got_mn[fin] - mn32[fin]).max() < 2e-6
    return n_interior


def test_contact_query_small(dev, small_assets):
    n = check_query(small_assets, dev, batch=5, seed=3, use_segments=True)
    assert n > 0
    check_query(small_assets, dev, batch=2, seed=4, use_segments=False)


def test_contact_query_matches_reference_golden(dev, small_assets):
    g = golden('contact_fitting_loss.npz')
    topo = make_topology(small_assets, dev, geothres=float(g['geothres']))
    verts = torch.tensor(g['thres02_seg/verts'], device=dev)
    out = topo.contact_query(verts, use_segments=True)
    assert np.abs(out['winding'].cpu().numpy() - g['winding']).max() < 2e-5
    assert np.array_equal(out['argmin'].cpu().numpy(), g['argmin'])


def test_segment_exterior_matches_reference_golden(dev, small_assets):
    s = golden('segments.npz')
    topo = make_topology(small_assets, dev)
    flags, _ = topo.segment_exterior(torch.tensor(s['verts'], device=dev))
    for name, f in zip(topo.segment_names, flags):
        for b in range(3):
            assert np.array_equal(f[b].cpu().numpy(), s['ext/%s/%d' % (name, b)]), (name, b)


def test_segmentation_mirror_api(dev, small_assets):
    from tuch_b200.utils.segmentation import BatchBodySegment
    s = golden('segments.npz')
    faces = torch.tensor(small_assets['model']['faces'], device=dev)
    bbs = BatchBodySegment(list(small_assets['segs'].keys()), faces, segment_data=small_assets['segs'])
    v = torch.tensor(s['verts'], device=dev)
    for name in bbs.names:
        seg = bbs.segmentation[name]
        assert np.array_equal(seg.segment_vidx, s['vidx/' + name])
        assert np.array_equal(seg.segment_faces.cpu().numpy(), s['faces/' + name])
        assert np.array_equal(seg.has_self_isect(v[[1]]).cpu().numpy(), s['ext/%s/1' % name])
    for name, e in zip(bbs.names, bbs.batch_has_self_isec(v[[2]])):
        assert np.array_equal(e.cpu().numpy(), s['ext/%s/2' % name])


def test_region_min_matches_oracle_and_golden(dev, small_assets):
    from oracle import clib
    c = golden('contact_from_verts.npz')
    topo = make_topology(small_assets, dev)
    verts = torch.tensor(c['verts'], device=dev)
    mn, ai, aj = topo.region_min(verts, masked=False)
    assert np.abs(mn.cpu().numpy() - c['value']).max() < 2e-6          # train_module.contact_from_verts
    geomask = small_assets['geo'] > 0.3
    reg = small_assets['regions']
    mn, ai, aj = topo.region_min(verts, masked=True)
    for b in range(verts.shape[0]):
        for k, (ra, rb) in enumerate(reg['classes']):
            ref, pa, pb = clib.region_min(c['verts'][b], geomask, reg['csig'][ra], reg['csig'][rb])
            got = float(mn[b, k])
            if np.isinf(ref):
                assert np.isinf(got)
                assert int(ai[b, k]) == reg['csig'][ra][0] and int(aj[b, k]) == reg['csig'][rb][0]
            else:
                assert abs(got - ref) < 2e-6
                assert geomask[int(ai[b, k]), int(aj[b, k])]


def test_contact_query_full_size(dev, full_assets):
    """SMPL-sized mesh (V=6890, F=13776): parity against the oracle on 2 bodies + size-independent
    properties on a larger batch."""
    n = check_query(full_assets, dev, batch=2, seed=11, use_segments=True)
    assert n > 0
    topo = make_topology(full_assets, dev, segments=False, regions=False)
    verts = torch.tensor(posed_verts(full_assets, 6, seed=12), device=dev)
    a = topo.contact_query(verts, use_segments=False)
    # batch invariance: a sub-batch gives identical nearest vertices (the winding sum only changes
    # its F-split summation order with the batch size) and the call is deterministic
    b = topo.contact_query(verts[2:5].contiguous(), use_segments=False)
    for k in ('argmin', 'min_sq'):
        assert torch.equal(a[k][2:5], b[k])
    assert (a['winding'][2:5] - b['winding']).abs().max() < 5e-6
    a2 = topo.contact_query(verts, use_segments=False)
    for k in ('argmin', 'min_sq', 'winding', 'exterior'):
        assert torch.equal(a[k], a2[k])
    # rigid-motion invariance of the flags and (near-)invariance of the winding number
    R = torch.tensor([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]], device=verts.device)
    moved = verts @ R.T + torch.tensor([0.3, -0.2, 0.5], device=verts.device)
    c = topo.contact_query(moved.contiguous(), use_segments=False)
    assert (a['winding'] - c['winding']).abs().max() < 5e-5
    safe = (a['winding'] - 0.99).abs() > 1e-3
    assert torch.equal(a['exterior'][safe], c['exterior'][safe])
    # the template (unposed, no self-contact) has no interior vertex
    t = torch.tensor(full_assets['model']['v_template'], device=verts.device)[None]
    assert bool(topo.contact_query(t, use_segments=False)['exterior'].all())


def _fast_vs_exact(assets, dev, batch, seed, template):
    ex = make_topology(assets, dev, segments=False, regions=False)
    fa = make_topology(assets, dev, segments=False, regions=False, exact=False, template=template)
    verts = torch.tensor(posed_verts(assets, batch, seed), device=dev)
    a = ex.contact_query(verts, use_segments=False, want_nearest=False)
    b = fa.contact_query(verts, use_segments=False, want_nearest=False)
    stats = fa.cluster_stats()
    assert stats['leaves'] >= len(assets['model']['faces']) // stats['leaf_faces'] and stats['mids'] >= 1 and stats['tops'] >= 1
    err = (a['winding'] - b['winding']).abs()
    print('hierarchical winding: max |w_fast - w_exact| = %.2e' % float(err.max()))
    assert float(err.max()) < FAR_FIELD_TOL, float(err.max())
    # every query the far field could misclassify was re-evaluated exactly
    band = (b['winding'] - 0.99).abs() < WC_MARGIN
    if bool(band.any()):
        assert float(err[band].max()) < 2e-5
    away = (a['winding'] - 0.99).abs() > 1e-4
    assert torch.equal(a['exterior'][away], b['exterior'][away])
    assert int((~a['exterior']).sum()) > 0
    # deterministic, and invariant to the batch composition
    b2 = fa.contact_query(verts, use_segments=False, want_nearest=False)
    assert torch.equal(b['winding'], b2['winding'])
    c = fa.contact_query(verts[1:3].contiguous(), use_segments=False, want_nearest=False)
    assert (c['winding'] - b['winding'][1:3]).abs().max() < 5e-6
    return float(err.max())


def test_hierarchical_winding_full_size(dev, full_assets, full_assets_uv):
    """Far-field winding numbers (clusters.cu) against the all-faces sum at SMPL size, on the SMPL-like
    lattice body and on the sliver-triangle UV body, with the hierarchy built from the template and
    lazily from the first posed body."""
    _fast_vs_exact(full_assets, dev, batch=6, seed=21, template=True)
    _fast_vs_exact(full_assets, dev, batch=3, seed=22, template=False)
    _fast_vs_exact(full_assets_uv, dev, batch=4, seed=23, template=True)


@pytest.mark.parametrize('radius', [0.02, 0.05, 0.0])
def test_contact_query_within_radius(dev, full_assets, radius):
    """tuch_contact_query_within: the nearest vertex only where losses.py:96-103 reads it.  Interior vertices and
    vertices with an allowed vertex within the radius get exactly the unlimited answer, the rest (-1, inf)."""
    from oracle import lbs as olbs
    from tuch_b200 import synthetic as syn
    tm = olbs.to_torch_model(full_assets['model'])
    pose = torch.tensor(np.concatenate([syn.fold_arms_pose(3, seed=70 + k, fold=0.7 + 0.1 * k) for k in range(3)]))
    betas = torch.tensor(np.random.default_rng(70).normal(0, 0.5, size=(9, 10)).astype(np.float32))
    verts = olbs.smpl_forward(tm, betas, pose[:, 3:], pose[:, :3])[0].to(dev).contiguous()
    topo = make_topology(full_assets, dev, regions=False, exact=False)
    full = topo.contact_query(verts, use_segments=True)
    lim = topo.contact_query(verts, use_segments=True, within=radius)
    assert torch.equal(full['exterior'], lim['exterior']) and torch.equal(full['winding'], lim['winding'])
    must = (full['winding'] > 0.99) | (full['min_sq'] <= radius * radius)
    assert int(must.sum()) > 100 and int((~must).sum()) > 10000
    assert torch.equal(lim['argmin'][must], full['argmin'][must]) and torch.equal(lim['min_sq'][must], full['min_sq'][must])
    rest_ok = (lim['argmin'] == -1) & torch.isinf(lim['min_sq']) | (lim['argmin'] == full['argmin']) & (lim['min_sq'] == full['min_sq'])
    assert bool(rest_ok.all())
    assert int((lim['argmin'][~must] == -1).sum()) > 0.9 * int((~must).sum())
    # the contact term over either answer is the same number and the same gradient
    from tuch_b200 import ops
    if radius > 0:
        ga, gb = torch.zeros_like(verts), torch.zeros_like(verts)
        la, _ = ops.contact_loss(verts, full['argmin'], full['exterior'], radius, g_points=ga)
        lb, _ = ops.contact_loss(verts, lim['argmin'], lim['exterior'], radius, g_points=gb)
        assert torch.equal(la, lb) and torch.equal(ga, gb) and float(ga.abs().max()) > 0


def test_group_nodes_from_shifted_child_moments(dev, full_assets, full_assets_uv):
    """The pack kernel forms mid and top nodes from their children's moments, re-expressed about the parent's
    centre (exact identities, clusters.cu add_shifted), instead of a second and third pass over the faces:
    compare every node record with the one computed straight from the node's own faces."""
    from oracle import lbs as olbs
    from tuch_b200 import synthetic as syn
    for assets, seed in ((full_assets, 61), (full_assets_uv, 62)):
        topo = make_topology(assets, dev, segments=False, regions=False, exact=False)
        tm = olbs.to_torch_model(assets['model'])
        pose = torch.tensor(syn.fold_arms_pose(5, seed=seed, fold=0.8))
        betas = torch.tensor(np.random.default_rng(seed).normal(0, 0.7, size=(5, 10)).astype(np.float32))
        verts = olbs.smpl_forward(tm, betas, pose[:, 3:], pose[:, :3])[0].to(dev).contiguous()
        st = topo.cluster_stats()
        a = topo.pack_nodes(verts, direct=True).double()
        b = topo.pack_nodes(verts, direct=False).double()
        groups = st['tops'] + st['mids']
        assert torch.equal(a[:, groups:], b[:, groups:])                     # leaves: the same arithmetic
        ga, gb = a[:, :groups], b[:, :groups]
        assert (ga[..., :3] - gb[..., :3]).abs().max() < 2e-6                # centres
        assert ((ga[..., 3] - gb[..., 3]).abs() / ga[..., 3]).max() < 1e-4   # squared opening radius
        # moments: zeroth .. third order scale with area x radius^k; compare per record against its largest entry
        # of the same order (cancellation makes single entries arbitrarily small)
        for lo, hi in ((4, 8), (8, 14), (14, 17), (17, 27)):
            scale = ga[..., lo:hi].abs().amax(-1, keepdim=True) + 1e-12
            assert ((ga[..., lo:hi] - gb[..., lo:hi]).abs() / scale).max() < 2e-3, (lo, hi)


def test_hierarchical_winding_touching_contact(dev, full_assets):
    """SMPLify-DC converges to touching contact, where vertices sit right at the 0.99 threshold: sweep
    the arm fold so that vertices cross the torso surface in small steps and compare the flags with the
    all-faces sum."""
    from oracle import lbs as olbs
    from tuch_b200 import synthetic as syn
    ex = make_topology(full_assets, dev, segments=False, regions=False)
    fa = make_topology(full_assets, dev, segments=False, regions=False, exact=False)
    tm = olbs.to_torch_model(full_assets['model'])
    folds = np.linspace(0.55, 1.0, 12)
    pose = torch.tensor(np.concatenate([syn.fold_arms_pose(1, seed=24, fold=float(f)) for f in folds]))
    verts = olbs.smpl_forward(tm, torch.zeros(len(folds), 10), pose[:, 3:], pose[:, :3])[0].to(dev).contiguous()
    a = ex.contact_query(verts, use_segments=False, want_nearest=False)
    b = fa.contact_query(verts, use_segments=False, want_nearest=False)
    away = (a['winding'] - 0.99).abs() > 1e-4
    assert torch.equal(a['exterior'][away], b['exterior'][away])
    n_int = (~a['exterior']).sum(1)
    assert int(n_int.min()) != int(n_int.max())          # the sweep really crosses the surface


def test_contact_query_batch64_properties(dev, full_assets):
    """Size-independent properties on a bench-sized batch (64 SMPL-sized bodies): the product path
    (hierarchical winding + pruned nearest) against the all-faces / dense kernels on every body,
    determinism, and invariance to where a body sits in the batch."""
    from oracle import lbs as olbs
    from tuch_b200 import synthetic as syn
    tm = olbs.to_torch_model(full_assets['model'])
    B = 64
    pose = torch.tensor(np.concatenate([syn.fold_arms_pose(8, seed=40 + k, fold=0.6 + 0.06 * k) for k in range(8)]))
    betas = torch.tensor(np.random.default_rng(40).normal(0, 0.5, size=(B, 10)).astype(np.float32))
    verts = olbs.smpl_forward(tm, betas, pose[:, 3:], pose[:, :3])[0].to(dev).contiguous()
    ex = make_topology(full_assets, dev, regions=False)
    fa = make_topology(full_assets, dev, regions=False, exact=False)
    a = ex.contact_query(verts, use_segments=True)
    b = fa.contact_query(verts, use_segments=True)
    assert torch.equal(a['argmin'], b['argmin']) and torch.equal(a['min_sq'], b['min_sq'])
    assert (a['winding'] - b['winding']).abs().max() < FAR_FIELD_TOL
    away = (a['winding'] - 0.99).abs() > 1e-4
    assert torch.equal(a['exterior'][away], b['exterior'][away])
    assert int((~b['exterior']).sum()) > 1000
    b2 = fa.contact_query(verts, use_segments=True)
    for k in ('argmin', 'min_sq', 'winding', 'exterior'):
        assert torch.equal(b[k], b2[k])
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).to(dev)
    c = fa.contact_query(verts[perm].contiguous(), use_segments=True)
    for k in ('argmin', 'min_sq', 'exterior'):
        assert torch.equal(b[k][perm], c[k])
    assert (b['winding'][perm] - c['winding']).abs().max() < 5e-6
