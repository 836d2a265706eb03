"""Stages the few UNMODIFIED reference files that bench.py's `ref_gpu` leg executes on the GPU box into
baseline/_ref/ (git-ignored: the reference's licence forbids redistribution, so they never enter the history; they
travel to the GPU box with the gpurun snapshot like the built .so files).  Run in the build container, where
/root/reference exists -- __graft_entry__.build() calls it.  On a box without the reference tree and without a
staged copy, bench.py's `ref_gpu` leg falls back to the restatement in tests/test_reference_chain_gpu.py and
says so in its `source` field.

    python scripts/stage_reference.py     ->   baseline/_ref/{configs/config.py, tuch/utils/{contact,geometry}.py,
                                                              tuch/smplify/{losses,prior}.py,
                                                              tuch/train/{train_module,fits_dict}.py}
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
DST = os.path.join(ROOT, 'baseline', '_ref')
FILES = ['configs/config.py', 'tuch/utils/contact.py', 'tuch/utils/geometry.py', 'tuch/smplify/losses.py',
         'tuch/smplify/prior.py',
         # the reference's CALLER code for the drop-in acceptance test (tests/test_dropin_gpu.py): its own train step
         # and fits store, run unchanged against the aliased tuch_b200 modules
         'tuch/train/train_module.py', 'tuch/train/fits_dict.py']


def stage(verbose=True):
    if not os.path.isdir(REF):
        return False
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    if verbose:
        print('staged %d reference files into %s' % (len(FILES), DST))
    return True


def staged():
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


if __name__ == '__main__':
    sys.exit(0 if stage() else 1)
