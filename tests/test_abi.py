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
