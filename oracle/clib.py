"""TEST INFRASTRUCTURE -- ctypes binding of oracle/contact_oracle.c (built by oracle/Makefile)."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'liboracle.so')
_lib = None


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ('contact_oracle.c', 'contact_oracle_impl.h')]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src)):
        return _SO
    subprocess.check_call(['make', '-C', _HERE, '-B'], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def num_threads():
    return int(lib().oracle_num_threads())


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return '_f32'
    if dtype == np.float64:
        return '_f64'
    raise TypeError(dtype)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _prep(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def solid_angles(points, tris, dtype=np.float32):
    """points[Q,3], tris[F,3,3] -> [Q,F]  (tuch/utils/contact.py:49-109)."""
    p, t = _prep(points, dtype), _prep(tris, dtype).reshape(-1, 9)
    out = np.empty((len(p), len(t)), dtype)
    getattr(lib(), 'oracle_solid_angles' + _sfx(dtype))(_p(p), _p(t), len(p), len(t), _p(out))
    return out


def winding_numbers(points, tris, dtype=np.float32):
    """points[Q,3], tris[F,3,3] -> [Q]  (tuch/utils/contact.py:112-147)."""
    p, t = _prep(points, dtype), _prep(tris, dtype).reshape(-1, 9)
    out = np.empty(len(p), dtype)
    getattr(lib(), 'oracle_winding_numbers' + _sfx(dtype))(_p(p), _p(t), len(p), len(t), _p(out))
    return out


def pairwise_dist(x, y, squared=True, dtype=np.float32):
    """x[Nx,3], y[Ny,3] -> [Nx,Ny]  (tuch/utils/contact.py:23-47, one batch element)."""
    x, y = _prep(x, dtype), _prep(y, dtype)
    out = np.empty((len(x), len(y)), dtype)
    getattr(lib(), 'oracle_pairwise_dist' + _sfx(dtype))(_p(x), _p(y), len(x), len(y), int(squared), _p(out))
    return out


def masked_nearest(v, geomask, dtype=np.float32):
    """v[V,3], geomask[V,V] bool -> (argmin[V] int32, min[V])  (tuch/smplify/losses.py:92-93)."""
    v = _prep(v, dtype)
    m = np.ascontiguousarray(geomask, dtype=np.uint8)
    V = len(v)
    assert m.shape == (V, V)
    am = np.empty(V, np.int32)
    mn = np.empty(V, dtype)
    getattr(lib(), 'oracle_masked_nearest' + _sfx(dtype))(_p(v), _p(m), V, _p(am), _p(mn))
    return am, mn


def region_min(v, geomask, ids_a, ids_b, dtype=np.float32):
    """min of the (optionally masked) squared distances over ids_a x ids_b -> (min, a_pos, b_pos)
    (tuch/smplify/losses.py:113-116; unmasked: tuch/train/train_module.py:83-90)."""
    v = _prep(v, dtype)
    a = np.ascontiguousarray(ids_a, dtype=np.int32)
    b = np.ascontiguousarray(ids_b, dtype=np.int32)
    m = None if geomask is None else np.ascontiguousarray(geomask, dtype=np.uint8)
    mn = np.empty(1, dtype)
    ia = np.empty(1, np.int32)
    ib = np.empty(1, np.int32)
    getattr(lib(), 'oracle_region_min' + _sfx(dtype))(
        _p(v), None if m is None else _p(m), len(v), _p(a), len(a), _p(b), len(b), _p(mn), _p(ia), _p(ib))
    return mn[0], int(ia[0]), int(ib[0])
