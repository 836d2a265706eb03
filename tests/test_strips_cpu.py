"""Host-only checks of the triangle-strip stream the winding kernel consumes: every face is closed
exactly once, with the right orientation sign, and every 256-element tile is self-contained."""
import numpy as np

TILE = 256


def replay(vid, flag):
    """-> list of (a, b, c, negate) triangles the kernel would close, with tile-local state."""
    out = []
    for t0 in range(0, len(vid), TILE):
        a = b = None
        for e in range(t0, t0 + TILE):
            c = int(vid[e])
            if flag[e] & 1:
                assert a is not None and b is not None and a >= 0 and b >= 0 and c >= 0, e
                out.append((a, b, c, bool(flag[e] >> 31)))
            a, b = b, c
    return out


def check(faces):
    from tuch_b200 import ops
    vid, flag, n_strips = ops.strip_stream(faces)
    assert len(vid) % TILE == 0 and len(vid) == len(flag)
    tris = replay(vid, flag)
    assert len(tris) == len(faces)
    canon = lambda t: min((t[i], t[(i + 1) % 3], t[(i + 2) % 3]) for i in range(3))
    want = sorted(canon(tuple(int(x) for x in f)) for f in faces)
    got = sorted(canon((a, c, b) if neg else (a, b, c)) for a, b, c, neg in tris)
    assert got == want
    return len(vid), n_strips


def test_strip_stream_small_and_full():
    from tuch_b200 import synthetic as syn
    for rings, segs in ((10, 12), (84, 82)):
        m = syn.make_body_model(rings, segs, seed=0)
        L, n = check(m['faces'])
        assert L < 1.08 * len(m['faces']) + TILE, (L, len(m['faces']), n)


def test_strip_stream_irregular_meshes():
    rng = np.random.default_rng(0)
    # a single triangle, two triangles with opposite orientation, a random triangle soup and a fan
    check(np.array([[0, 1, 2]]))
    check(np.array([[0, 1, 2], [1, 0, 3]]))
    check(np.array([[0, 1, 2], [0, 1, 3]]))             # inconsistent orientation across the shared edge
    soup = np.array([rng.choice(40, size=3, replace=False) for _ in range(700)])
    check(soup)
    fan = np.array([[0, i, i + 1] for i in range(1, 600)])
    check(fan)
