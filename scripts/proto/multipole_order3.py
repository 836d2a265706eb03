"""Prototype (numpy, CPU): third-order far-field term for the GROUP nodes of the face hierarchy.

The hierarchical winding kernel expands the solid angle of a node about its centre p to second order in the
source offsets d = x - p.  With r = p - q and phi = 1/|r| the series is
    Omega(q) = sum_k (-1)^k / k! * D^{k+1} phi (r) . M_k ,   M_k = int n (x) d^(k) dA ,
and the next term (k = 3) needs the fully symmetrised fourth-rank moment S = sym(M_3) only, because D^4 phi is
symmetric and trace-free:
    D^4 phi . S = 105 S(r,r,r,r) / R^9 - 90 tr(S)(r,r) / R^7 + 9 tr tr(S) / R^5 .
This script measures, on the synthetic lattice body, the largest far-field error of a mid / top group seen
from the mesh vertices that would NOT open it, at opening radii beta = 2.0 and 2.5, for the second- and
third-order series.  Moments are integrated with a degree-5 triangle rule.
    python scripts/proto/multipole_order3.py
"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

from tuch_b200 import ops, synthetic as syn

# Dunavant degree-5 rule (7 points): barycentric coordinates and weights
_A1, _B1 = 0.059715871789770, 0.470142064105115
_A2, _B2 = 0.797426985353087, 0.101286507323456
BARY = np.array([[1 / 3, 1 / 3, 1 / 3], [_A1, _B1, _B1], [_B1, _A1, _B1], [_B1, _B1, _A1],
                 [_A2, _B2, _B2], [_B2, _A2, _B2], [_B2, _B2, _A2]])
WQ = np.array([0.225] + [0.132394152788506] * 3 + [0.125939180544827] * 3)


def exact_half_angles(q, tri):
    """Van Oosterom-Strackee, q[Q,3], tri[F,3,3] -> [Q,F] solid angles."""
    a, b, c = (tri[None, :, k, :] - q[:, None, :] for k in range(3))
    la, lb, lc = (np.linalg.norm(x, axis=-1) for x in (a, b, c))
    num = np.einsum('qfi,qfi->qf', a, np.cross(b, c))
    den = la * lb * lc + (a * b).sum(-1) * lc + (a * c).sum(-1) * lb + (b * c).sum(-1) * la
    return 2 * np.arctan2(num, den)


def moments(tri, p):
    """M0[3], M1[3,3], M2[3,3,3], M3[3,3,3,3] of the faces tri[F,3,3] about p."""
    n = 0.5 * np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])            # area-weighted normals
    x = np.einsum('gk,fkc->fgc', BARY, tri) - p                                 # [F,7,3] quadrature points
    w = WQ[None, :, None]
    M0 = n.sum(0)
    M1 = np.einsum('fi,fgj,g->ij', n, x, WQ)
    M2 = np.einsum('fi,fgj,fgk,g->ijk', n, x, x, WQ)
    M3 = np.einsum('fi,fgj,fgk,fgl,g->ijkl', n, x, x, x, WQ)
    return M0, M1, M2, M3


def dphi(r):
    """D^k phi(r), k = 1..4, phi = 1/|r|, for r[Q,3]."""
    R = np.linalg.norm(r, axis=-1)
    I = np.eye(3)
    D1 = -r / R[:, None] ** 3
    D2 = 3 * np.einsum('qi,qj->qij', r, r) / R[:, None, None] ** 5 - I / R[:, None, None] ** 3
    rr = np.einsum('qi,qj,qk->qijk', r, r, r)
    dr = (np.einsum('ij,qk->qijk', I, r) + np.einsum('ik,qj->qijk', I, r) + np.einsum('jk,qi->qijk', I, r))
    D3 = -15 * rr / R[:, None, None, None] ** 7 + 3 * dr / R[:, None, None, None] ** 5
    rrrr = np.einsum('qi,qj,qk,ql->qijkl', r, r, r, r)
    d_rr = sum(np.einsum('%s,q%s,q%s->qijkl' % (a + b, c, d), I, r, r)
               for (a, b, c, d) in [('i', 'j', 'k', 'l'), ('i', 'k', 'j', 'l'), ('i', 'l', 'j', 'k'),
                                    ('j', 'k', 'i', 'l'), ('j', 'l', 'i', 'k'), ('k', 'l', 'i', 'j')])
    dd = (np.einsum('ij,kl->ijkl', I, I) + np.einsum('ik,jl->ijkl', I, I) + np.einsum('il,jk->ijkl', I, I))
    R5 = R[:, None, None, None, None]
    D4 = 105 * rrrr / R5 ** 9 - 15 * d_rr / R5 ** 7 + 3 * dd[None] / R5 ** 5
    return D1, D2, D3, D4


def series(q, p, M):
    """Solid angle of the node from q[Q,3] to order 2 and order 3: n.g(x - q), g = -grad phi at (r + d)."""
    M0, M1, M2, M3 = M
    D1, D2, D3, D4 = dphi(p[None] - q)
    t0 = -np.einsum('qi,i->q', D1, M0)
    t1 = -np.einsum('qij,ij->q', D2, M1)
    t2 = -0.5 * np.einsum('qijk,ijk->q', D3, M2)
    t3 = -np.einsum('qijkl,ijkl->q', D4, M3) / 6.0
    return t0 + t1 + t2, t0 + t1 + t2 + t3


def main():
    model = syn.make_lattice_body_model(seed=0)
    verts, faces = model['v_template'].astype(np.float64), model['faces']
    tree = ops.cluster_tree(faces, verts)
    leaf, mid, top = tree['leaf_face'], tree['mid_off'], tree['top_off']
    groups = [('mid', leaf[mid[m]:mid[m + 1]]) for m in range(len(mid) - 1)]
    groups += [('top', leaf[mid[top[t]]:mid[top[t + 1]]]) for t in range(len(top) - 1)]
    worst = {(o, b): 0.0 for o in (2, 3) for b in (2.0, 2.5)}
    total = {(o, b): np.zeros(len(verts)) for o in (2, 3) for b in (2.0, 2.5)}
    for kind, rows in groups:
        fid = rows[rows >= 0]
        tri = verts[faces[fid]]
        area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
        p = (area[:, None] * tri.mean(1)).sum(0) / area.sum()
        R = np.linalg.norm(tri.reshape(-1, 3) - p, axis=1).max()
        d = np.linalg.norm(verts - p, axis=1)
        M = moments(tri, p)
        for beta in (2.0, 2.5):
            far = np.where(d >= beta * R)[0]
            if len(far) == 0:
                continue
            ex = exact_half_angles(verts[far], tri).sum(1)
            o2, o3 = series(verts[far], p, M)
            for order, val in ((2, o2), (3, o3)):
                err = np.abs(val - ex) / (4 * np.pi)
                worst[(order, beta)] = max(worst[(order, beta)], err.max())
                if kind == 'mid':                                  # one level's errors add up per query
                    total[(order, beta)][far] += err
    print('%d mids + %d tops on the lattice body (V=%d, F=%d)' % (len(mid) - 1, len(top) - 1, len(verts), len(faces)))
    for (order, beta), w in sorted(worst.items()):
        print('order %d, groups opened at %.1f radii: largest single-group error %.2e of a winding number, '
              'largest summed error over the mids of one query %.2e' % (order, beta, w, total[(order, beta)].max()))


if __name__ == '__main__':
    main()
