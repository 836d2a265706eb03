"""Host-only checks of the face-cluster hierarchy behind the hierarchical winding kernel
(tuch_b200/csrc/clusters.cu), and of the far-field expansion it evaluates, restated in numpy and held
against the oracle's exact solid angles."""
import numpy as np


def tree_for(model):
    from tuch_b200 import ops
    return ops.cluster_tree(model['faces'], model['v_template'])


def check_tree(model):
    faces = model['faces']
    F, V = len(faces), len(model['v_template'])
    t = tree_for(model)
    leaf, mid, top, vt = t['leaf_face'], t['mid_off'], t['top_off'], t['vtile']
    ids = leaf[leaf >= 0]
    assert len(ids) == F and np.array_equal(np.sort(ids), np.arange(F))            # a partition of the faces
    assert np.array_equal(np.sort(vt[vt >= 0]), np.arange(V))                      # a partition of the vertices
    assert mid[0] == 0 and mid[-1] == len(leaf) and np.all(np.diff(mid) >= 1) and np.all(np.diff(mid) <= 24)
    assert top[0] == 0 and top[-1] == len(mid) - 1 and np.all(np.diff(top) >= 1)
    for row in list(leaf) + list(vt):                                             # ascending ids, padding only at the end
        n = int((row >= 0).sum())
        assert n >= 1 and np.all(row[:n] >= 0) and np.all(row[n:] < 0) and np.all(np.diff(row[:n]) > 0)
    return t


def test_tree_small_and_full():
    from tuch_b200 import synthetic as syn
    check_tree(syn.make_body_model(10, 12, seed=0))
    t = check_tree(syn.make_lattice_body_model(seed=0))
    # near-minimal leaf count: the far field costs one evaluation per leaf
    assert len(t['leaf_face']) <= 1.15 * (13776 // 16 + 1) and len(t['vtile']) <= 1.15 * (6890 // 32 + 1)
    check_tree(syn.make_body_model(84, 82, seed=0))


def test_tree_disconnected_and_tiny():
    from tuch_b200 import ops
    # two separate tetrahedra and a lone triangle
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    verts = np.concatenate([v, v + 5, v[:3] + 9])
    tet = np.array([[0, 2, 1], [0, 1, 3], [1, 2, 3], [0, 3, 2]])
    faces = np.concatenate([tet, tet + 4, [[8, 9, 10]]])
    t = ops.cluster_tree(faces, verts)
    assert sorted(t['leaf_face'][t['leaf_face'] >= 0].tolist()) == list(range(9))
    assert sorted(t['vtile'][t['vtile'] >= 0].tolist()) == list(range(11))
    assert len(t['leaf_face']) == 1 and len(t['vtile']) == 1                       # small components share a leaf


WC_BETA, WC_MARGIN = 1.4, 0.10          # tuch_b200/csrc/clusters.h


def test_far_field_expansion_matches_exact_solid_angles():
    """The node record of cluster_pack_kernel, restated in fp64: centre, radius, M0, tr M1, sym M1, u and
    the cubic form.  Beyond WC_BETA radii its error against the exact cluster solid angles is a few 1e-3
    of a winding number in total."""
    import torch
    from oracle import clib, lbs as olbs
    from tuch_b200 import synthetic as syn
    model = syn.make_lattice_body_model(seed=0)
    faces = model['faces']
    t = tree_for(model)
    tm = olbs.to_torch_model(model)
    pose = torch.tensor(syn.fold_arms_pose(1, seed=5))
    v = olbs.smpl_forward(tm, torch.zeros(1, 10), pose[:, 3:], pose[:, :3])[0][0].numpy().astype(np.float64)
    tri = v[faces]
    A, B, C = tri[:, 0], tri[:, 1], tri[:, 2]
    av = 0.5 * np.cross(B - A, C - A)
    area = np.linalg.norm(av, axis=1)
    rng = np.random.default_rng(0)
    q = v[rng.choice(len(v), 400, replace=False)]
    sa = clib.solid_angles(q, tri, dtype=np.float64)                               # [Q,F], = 2 atan2(...)
    total_err = np.zeros(len(q))
    signed_err = np.zeros(len(q))
    for row in t['leaf_face']:
        fs = row[row >= 0]
        p = ((A[fs] + B[fs] + C[fs]) / 3 * area[fs, None]).sum(0) / area[fs].sum()
        R = np.sqrt(((tri[fs].reshape(-1, 3) - p) ** 2).sum(1).max())
        c = (A[fs] + B[fs] + C[fs]) / 3 - p
        mids = np.stack([(A[fs] + B[fs]) / 2, (B[fs] + C[fs]) / 2, (C[fs] + A[fs]) / 2], 1) - p
        S = np.einsum('tmj,tmk->tjk', mids, mids) / 3
        M0 = av[fs].sum(0)
        M1 = np.einsum('ti,tj->ij', av[fs], c)
        M2 = np.einsum('ti,tjk->ijk', av[fs], S)
        u = 2 * np.einsum('iik->k', M2) + np.einsum('kjj->k', M2)
        r = p[None] - q
        d = np.linalg.norm(r, axis=1)
        omega = (r @ M0 + np.trace(M1)) / d ** 3 - (3 * np.einsum('qi,ij,qj->q', r, M1, r) + 1.5 * r @ u) / d ** 5 \
            + 7.5 * np.einsum('ijk,qi,qj,qk->q', M2, r, r, r) / d ** 7
        far = d > WC_BETA * R
        total_err += np.where(far, np.abs(omega - sa[:, fs].sum(1)), 0.0) / (4 * np.pi)
        signed_err += np.where(far, omega - sa[:, fs].sum(1), 0.0) / (4 * np.pi)
    # even if every leaf's error had the same sign (no cancellation at all) the total stays below HALF the
    # re-evaluation margin (WC_MARGIN); the actual, signed error stays below a quarter of it
    print('far-field bound: sum |err| %.3e, |sum err| %.3e' % (total_err.max(), np.abs(signed_err).max()))
    assert total_err.max() < 0.5 * WC_MARGIN, total_err.max()
    assert np.abs(signed_err).max() < 0.25 * WC_MARGIN, np.abs(signed_err).max()


def test_refined_tree_is_a_valid_partition_with_rounder_leaves():
    """TUCH_TREE_REFINE (off by default, read once per process -> a subprocess here): the refined hierarchy is
    still a partition with ascending ids, keeps the leaf / group counts, and its leaves are rounder."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, json
sys.path.insert(0, %r)
import numpy as np
from tuch_b200 import ops, synthetic as syn
m = syn.make_lattice_body_model(seed=0)
v, f = m['v_template'].astype(np.float64), m['faces']
t = ops.cluster_tree(f, v)
leaf, vt = t['leaf_face'], t['vtile']
assert np.array_equal(np.sort(leaf[leaf >= 0]), np.arange(len(f))) and np.array_equal(np.sort(vt[vt >= 0]), np.arange(len(v)))
for row in list(leaf) + list(vt):
    n = int((row >= 0).sum())
    assert n >= 1 and np.all(row[:n] >= 0) and np.all(row[n:] < 0) and np.all(np.diff(row[:n]) > 0)
tri = v[f]
r3 = 0.0
for row in leaf:
    c = tri[row[row >= 0]].reshape(-1, 3)
    r3 += np.linalg.norm(c - c.mean(0), axis=1).max() ** 3
print(json.dumps(dict(K=len(leaf), NM=len(t['mid_off']) - 1, NT=len(t['top_off']) - 1, T=len(vt), r3=r3,
                      mid_off=t['mid_off'].tolist())))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    for tag, sweeps in (('plain', '0'), ('refined', '3')):
        env = dict(os.environ, TUCH_TREE_REFINE=sweeps)
        r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        import json
        out[tag] = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ('K', 'NM', 'NT', 'T', 'mid_off'):
        assert out['plain'][k] == out['refined'][k], k
    assert out['refined']['r3'] < 0.85 * out['plain']['r3'], (out['refined']['r3'], out['plain']['r3'])


def _raw_moments(tris, p):
    """The 23 raw sums NodeMoments::add accumulates (tuch_b200/csrc/clusters.cu), in float64, about centre p."""
    out = np.zeros(23)
    for t in tris:
        a, b, g = t - p
        n = 0.5 * np.cross(b - a, g - a)
        h = (a + b + g) / 3
        S = sum(np.outer(x, x) for x in (0.5 * (a + b), 0.5 * (b + g), 0.5 * (g + a))) / 3
        out[0:3] += n
        out[3] += n @ h
        out[4:7] += n * h
        out[7] += n[0] * h[1] + n[1] * h[0]
        out[8] += n[0] * h[2] + n[2] * h[0]
        out[9] += n[1] * h[2] + n[2] * h[1]
        out[10:13] += 2 * S @ n + n * np.trace(S)
        nx, ny, nz = n
        sxx, syy, szz, sxy, sxz, syz = S[0, 0], S[1, 1], S[2, 2], S[0, 1], S[0, 2], S[1, 2]
        out[13:16] += [nx * sxx, ny * syy, nz * szz]
        out[16] += 2 * nx * sxy + ny * sxx
        out[17] += 2 * nx * sxz + nz * sxx
        out[18] += 2 * ny * sxy + nx * syy
        out[19] += 2 * ny * syz + nz * syy
        out[20] += 2 * nz * sxz + nx * szz
        out[21] += 2 * nz * syz + ny * szz
        out[22] += 2 * (nx * syz + ny * sxz + nz * sxy)
    return out


def _shifted(s, d):
    """clusters.cu add_shifted: the sums of a child about its own centre re-expressed about a centre d away."""
    dx, dy, dz = d
    m0x, m0y, m0z, tr, qxx, qyy, qzz, qxy, qxz, qyz = s[:10]
    dm, d2 = d @ s[:3], d @ d
    trd = tr + dm
    o = np.zeros(23)
    o[0:3] = s[0:3]
    o[3] = trd
    o[4:7] = [qxx + m0x * dx, qyy + m0y * dy, qzz + m0z * dz]
    o[7:10] = [qxy + m0x * dy + m0y * dx, qxz + m0x * dz + m0z * dx, qyz + m0y * dz + m0z * dy]
    o[10] = s[10] + 2 * (2 * qxx * dx + qxy * dy + qxz * dz) + 2 * dx * trd + d2 * m0x
    o[11] = s[11] + 2 * (qxy * dx + 2 * qyy * dy + qyz * dz) + 2 * dy * trd + d2 * m0y
    o[12] = s[12] + 2 * (qxz * dx + qyz * dy + 2 * qzz * dz) + 2 * dz * trd + d2 * m0z
    o[13] = s[13] + 2 * dx * qxx + dx * dx * m0x
    o[14] = s[14] + 2 * dy * qyy + dy * dy * m0y
    o[15] = s[15] + 2 * dz * qzz + dz * dz * m0z
    o[16] = s[16] + 2 * (dx * qxy + dy * qxx) + dx * dx * m0y + 2 * dx * dy * m0x
    o[17] = s[17] + 2 * (dx * qxz + dz * qxx) + dx * dx * m0z + 2 * dx * dz * m0x
    o[18] = s[18] + 2 * (dy * qxy + dx * qyy) + dy * dy * m0x + 2 * dx * dy * m0y
    o[19] = s[19] + 2 * (dy * qyz + dz * qyy) + dy * dy * m0z + 2 * dy * dz * m0y
    o[20] = s[20] + 2 * (dz * qxz + dx * qzz) + dz * dz * m0x + 2 * dx * dz * m0z
    o[21] = s[21] + 2 * (dz * qyz + dy * qzz) + dz * dz * m0y + 2 * dy * dz * m0z
    o[22] = s[22] + 2 * (dx * qyz + dy * qxz + dz * qxy) + 2 * (dx * dy * m0z + dx * dz * m0y + dy * dz * m0x)
    return o


def test_child_moments_shift_to_the_parent_centre_exactly():
    """The pack kernel forms group nodes from their children's moments (add_shifted).  The identities are exact:
    in float64 the shifted sums of two child clusters add up to the sums taken straight about the parent's centre."""
    rng = np.random.default_rng(5)
    tris_a, tris_b = rng.normal(size=(7, 3, 3)), rng.normal(size=(5, 3, 3)) + 2.0
    pa, pb, p = tris_a.mean((0, 1)), tris_b.mean((0, 1)), rng.normal(size=3)
    direct = _raw_moments(np.concatenate([tris_a, tris_b]), p)
    combined = _shifted(_raw_moments(tris_a, pa), pa - p) + _shifted(_raw_moments(tris_b, pb), pb - p)
    assert np.abs(direct - combined).max() < 1e-11 * np.abs(direct).max()
