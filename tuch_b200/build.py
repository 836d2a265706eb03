"""Builds tuch_b200/lib/libtuch_b200.so (sm_100a only) with nvcc.  No GPU needed to compile."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'lib', 'libtuch_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC,-fvisibility=hidden']


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def headers():
    inc = os.path.join(os.path.dirname(HERE), 'include')
    return (sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh')))
            + sorted(os.path.join(inc, f) for f in os.listdir(inc) if f.endswith('.h')))


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(f) <= t for f in sources() + headers())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found; cannot build libtuch_b200.so')
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB] + sources()
    env = dict(os.environ)
    env.pop('CC', None)       # the image's CC points at a gcc wrapper nvcc should not use as host compiler
    env.pop('CXX', None)
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
