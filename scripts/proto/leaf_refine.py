"""Prototype (numpy, CPU): how much a pairwise re-bisection of neighbouring face leaves shrinks the near set of
the hierarchical winding kernel.  Starting from the leaves of clusters.cu (PCA bisection), every pair of leaves
that share a mesh edge is pooled and re-split -- along the best of 12 directions, sizes kept <= 16 -- whenever
that lowers R_a^3 + R_b^3 (R = bounding radius about the area-weighted centre); a few sweeps.  Reported: the
(query vertex, leaf) pairs inside the 2 R opening radius on the template pose, before and after.
    python scripts/proto/leaf_refine.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

from tuch_b200 import ops, synthetic as syn

LEAF = 16


def main():
    model = syn.make_lattice_body_model(seed=0)
    v, f = model['v_template'].astype(np.float64), model['faces']
    tri = v[f]
    area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    cen = tri.mean(1)
    leaves = [row[row >= 0].tolist() for row in ops.cluster_tree(f, v)['leaf_face']]

    def sphere(ids):
        a = area[ids]
        p = (a[:, None] * cen[ids]).sum(0) / a.sum()
        return p, np.linalg.norm(tri[ids].reshape(-1, 3) - p, axis=1).max()

    def near_pairs(ls):
        n = 0
        for ids in ls:
            p, r = sphere(np.asarray(ids))
            n += int((np.linalg.norm(v - p, axis=1) < 2.0 * r).sum())
        return n

    def stats(ls, tag):
        r = np.array([sphere(np.asarray(ids))[1] for ids in ls])
        ideal = np.array([np.sqrt(area[ids].sum() / np.pi) for ids in ls])
        print('%-8s %d leaves, R/disc median %.2f p90 %.2f, sum R^3 %.5f, near (vertex, leaf) pairs %d'
              % (tag, len(ls), np.median(r / ideal), np.percentile(r / ideal, 90), (r ** 3).sum(), near_pairs(ls)))

    stats(leaves, 'before')
    # leaf adjacency through shared mesh edges
    edge_owner = {}
    owner = np.empty(len(f), np.int64)
    for li, ids in enumerate(leaves):
        owner[ids] = li
    for fi, (a, b, c) in enumerate(f):
        for e in ((a, b), (b, c), (c, a)):
            edge_owner.setdefault((min(e), max(e)), []).append(fi)
    rng = np.random.default_rng(0)
    dirs = rng.normal(size=(12, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    for sweep in range(4):
        pairs = {(min(owner[x], owner[y]), max(owner[x], owner[y])) for fs in edge_owner.values() if len(fs) == 2
                 for x, y in [fs] if owner[x] != owner[y]}
        improved = 0
        for a, b in sorted(pairs):
            ia, ib = np.asarray(leaves[a]), np.asarray(leaves[b])
            pool = np.concatenate([ia, ib])
            cost = sphere(ia)[1] ** 3 + sphere(ib)[1] ** 3
            best = None
            lo = max(len(pool) - LEAF, 1)
            for d in dirs:
                order = pool[np.argsort(cen[pool] @ d)]
                for k in range(lo, min(LEAF, len(pool) - 1) + 1):
                    c = sphere(order[:k])[1] ** 3 + sphere(order[k:])[1] ** 3
                    if c < cost * 0.999:
                        cost, best = c, (order[:k].tolist(), order[k:].tolist())
            if best is not None:
                leaves[a], leaves[b] = best
                owner[best[0]] = a
                owner[best[1]] = b
                improved += 1
        stats(leaves, 'sweep %d' % (sweep + 1))
        if improved == 0:
            break


if __name__ == '__main__':
    main()
