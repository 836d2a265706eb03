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
