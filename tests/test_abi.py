"""CPU-side checks of the drop-in boundary: the shared library builds, loads and exports every
symbol include/tuch_b200.h declares; the product path fails loudly without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'tuch_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(tuch_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_and_exports_every_declared_symbol():
    from tuch_b200 import build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.tuch_abi_version() == 1


def test_every_exported_symbol_is_declared():
    import subprocess
    from tuch_b200 import build
    out = subprocess.check_output(['nm', '-D', '--defined-only', build.build()]).decode()
    exported = sorted(l.split()[-1] for l in out.splitlines() if ' T ' in l and 'tuch_' in l)
    assert exported == declared_symbols()


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback():
    from tuch_b200 import ops
    from tuch_b200.utils import contact
    x = torch.zeros(1, 4, 3)
    with pytest.raises(ops.TuchError):
        contact.batch_pairwise_dist(x, x)
    with pytest.raises(ops.TuchError):
        contact.winding_numbers(x, torch.zeros(1, 2, 3, 3))
    with pytest.raises(ops.TuchError):
        ops.Topology([[0, 1, 2]], 3, 'cpu')


def test_host_entry_points_validate_their_arguments():
    """Error convention of the ABI on the entry points that need no device: non-zero return code and a
    message from tuch_last_error(); nothing is written on failure."""
    import numpy as np
    from tuch_b200 import build
    lib = ctypes.CDLL(build.build())
    lib.tuch_last_error.restype = ctypes.c_char_p
    i32p = ctypes.POINTER(ctypes.c_int32)
    faces = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    verts = np.zeros((4, 3), np.float32)
    n = [ctypes.c_int32(-7) for _ in range(4)]
    refs = [ctypes.byref(x) for x in n]
    fp, vp = faces.ctypes.data_as(ctypes.c_void_p), verts.ctypes.data_as(ctypes.c_void_p)
    # a face index out of range
    bad = faces.copy()
    bad[1, 2] = 9
    rc = lib.tuch_cluster_tree_host(bad.ctypes.data_as(ctypes.c_void_p), 2, 4, vp, None, 0, None, 0, None, 0, None, 0, *refs)
    assert rc != 0 and b'out of range' in lib.tuch_last_error() and n[0].value == -7
    # null mesh
    assert lib.tuch_cluster_tree_host(None, 2, 4, vp, None, 0, None, 0, None, 0, None, 0, *refs) != 0
    # output buffers that are too small
    assert lib.tuch_cluster_tree_host(fp, 2, 4, vp, None, 0, None, 0, None, 0, None, 0, *refs) == 0
    k = n[0].value
    assert k >= 1
    small = np.zeros((k, 16), np.int32)
    rc = lib.tuch_cluster_tree_host(fp, 2, 4, vp, small.ctypes.data_as(ctypes.c_void_p), k - 1, None, 0, None, 0, None, 0, *refs)
    assert rc != 0 and b'capacity' in lib.tuch_last_error()
    # the strip builder rejects an empty face list
    L = ctypes.c_int32(0)
    assert lib.tuch_strip_stream_host(fp, 0, None, None, 0, ctypes.byref(L), None) != 0
