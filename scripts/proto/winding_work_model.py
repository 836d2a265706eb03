"""Prototype (numpy, CPU): instruction-count model of winding_cluster_kernel for a given face / vertex hierarchy.

Replays the kernel's traversal on the template pose of the synthetic lattice body -- one warp per vertex tile,
tops and mids opened for the whole warp as soon as one lane is inside the opening radius, leaves near or far per
lane, the near pass two queries per step -- and counts node tests, far-field evaluations and near steps.  The
weights are the SASS instruction counts of those pieces.  Use it to compare tree variants
offline, e.g.
    python scripts/proto/winding_work_model.py [leaves per sub-group, default 0 = the shipped three levels]
    TUCH_TREE_REFINE=4 python scripts/proto/winding_work_model.py 2
A sub-group is a hypothetical fourth level: runs of N consecutive leaves inside a mid, opened like a group.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

from tuch_b200 import ops, synthetic as syn

BETA_LEAF, BETA_GROUP = 1.6, 2.0
W_TEST, W_FAR, W_NEAR_STEP = 15, 50, 84          # warp instructions per node test / far field / near step


def node_sphere(tri, area, cen, ids):
    a = area[ids]
    p = (a[:, None] * cen[ids]).sum(0) / a.sum()
    return p, np.linalg.norm(tri[ids].reshape(-1, 3) - p, axis=1).max()


def main():
    model = syn.make_lattice_body_model(seed=0)
    v, f = model['v_template'].astype(np.float64), model['faces']
    t = ops.cluster_tree(f, v)
    leaf, mid, top, vt = t['leaf_face'], t['mid_off'], t['top_off'], t['vtile']
    tri = v[f]
    area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    cen = tri.mean(1)
    leaf_ids = [row[row >= 0] for row in leaf]
    ls = [node_sphere(tri, area, cen, ids) for ids in leaf_ids]
    ms = [node_sphere(tri, area, cen, np.concatenate(leaf_ids[mid[m]:mid[m + 1]])) for m in range(len(mid) - 1)]
    ts = [node_sphere(tri, area, cen, np.concatenate(leaf_ids[mid[top[k]]:mid[top[k + 1]]])) for k in range(len(top) - 1)]
    sub = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    subs = {}
    if sub > 0:
        for m in range(len(mid) - 1):
            for l0 in range(mid[m], mid[m + 1], sub):
                l1 = min(l0 + sub, mid[m + 1])
                subs[l0] = (l1, node_sphere(tri, area, cen, np.concatenate(leaf_ids[l0:l1])))
    tests = far = near_steps = mids_open = 0
    for row in vt:
        q = v[row[row >= 0]]
        for k in range(len(top) - 1):
            tests += 1
            if not (np.linalg.norm(q - ts[k][0], axis=1) < BETA_GROUP * ts[k][1]).any():
                far += 1
                continue
            for m in range(top[k], top[k + 1]):
                tests += 1
                if not (np.linalg.norm(q - ms[m][0], axis=1) < BETA_GROUP * ms[m][1]).any():
                    far += 1
                    continue
                mids_open += 1
                l = mid[m]
                while l < mid[m + 1]:
                    l1 = l + 1
                    if sub > 0:
                        l1, (p, r) = subs[l]
                        tests += 1
                        if not (np.linalg.norm(q - p, axis=1) < BETA_GROUP * r).any():
                            far += 1
                            l = l1
                            continue
                    for ll in range(l, l1):
                        tests += 1
                        n_near = int((np.linalg.norm(q - ls[ll][0], axis=1) < BETA_LEAF * ls[ll][1]).sum())
                        if n_near < len(q):
                            far += 1
                        near_steps += (n_near + 1) // 2
                    l = l1
    n_warps = len(vt)
    total = W_TEST * tests + W_FAR * far + W_NEAR_STEP * near_steps
    print('TUCH_TREE_REFINE=%s, sub-groups of %d: per warp %.0f node tests, %.0f far fields, %.1f mids opened, %.0f near steps -> %.1f k '
          'warp instructions (tests %.0f %%, far %.0f %%, near %.0f %%)'
          % (os.environ.get('TUCH_TREE_REFINE', '0'), sub, tests / n_warps, far / n_warps, mids_open / n_warps,
             near_steps / n_warps, total / n_warps / 1e3, 100 * W_TEST * tests / total, 100 * W_FAR * far / total,
             100 * W_NEAR_STEP * near_steps / total))


if __name__ == '__main__':
    main()
