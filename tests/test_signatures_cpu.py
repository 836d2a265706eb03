"""Drop-in boundary: every mirror in tuch_b200/ takes the reference's parameters -- same names, same
order, same defaults -- as recorded from the reference sources by tests/golden/make_signatures.py
(SURVEY.md 8(b)).  Mirrors may append extra keyword parameters after the reference's own."""
import importlib
import inspect
import json
import os

import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'signatures.json')
MIRROR = {
    'tuch/utils/contact.py': 'tuch_b200.utils.contact',
    'tuch/utils/segmentation.py': 'tuch_b200.utils.segmentation',
    'tuch/utils/geometry.py': 'tuch_b200.utils.geometry',
    'tuch/models/smpl.py': 'tuch_b200.models.smpl',
    'tuch/smplify/prior.py': 'tuch_b200.smplify.prior',
    'tuch/smplify/losses.py': 'tuch_b200.smplify.losses',
    'tuch/smplify/smplifydc.py': 'tuch_b200.smplify.smplifydc',
    'tuch/train/loss.py': 'tuch_b200.train.loss',
    'tuch/train/fits_dict.py': 'tuch_b200.train.fits_dict',
    'tuch/train/train_module.py': 'tuch_b200.train.train_module',
    'tuch/eft/loss.py': 'tuch_b200.eft.loss',
}
# functions the mirrors expose at module level although the reference hangs them on a class that is out
# of scope as a whole (TUCH / EFTLoss: the trainer objects); `self` is dropped for those
MODULE_LEVEL = {'TUCH.contact_from_verts': 'contact_from_verts', 'EFTLoss.contact_loss': 'contact_loss'}
# constructor arguments the mirrors replace: the reference reads these from its un-shipped data tree
CTOR_FREE = {'MaxMixturePrior.__init__', 'RegressorLoss.__init__', 'BodySegment.__init__'}

with open(GOLDEN) as f:
    SIGS = json.load(f)


def _default_src(v):
    if v is inspect.Parameter.empty:
        return None
    import torch
    if isinstance(v, torch.device):
        return "torch.device('%s')" % v.type
    if v is torch.float32:
        return 'torch.float32'
    return repr(v)


@pytest.mark.parametrize('key', sorted(SIGS))
def test_mirror_signature(key):
    path, name = key.split(':')
    ref = SIGS[key]
    mod = importlib.import_module(MIRROR[path])
    ref_params, ref_defaults = list(ref['params']), list(ref['defaults'])
    if name in MODULE_LEVEL:
        obj = getattr(mod, MODULE_LEVEL[name])
        ref_params, ref_defaults = ref_params[1:], ref_defaults[1:]
    else:
        obj = mod
        for part in name.split('.'):
            obj = getattr(obj, part)
    got = list(inspect.signature(obj).parameters.values())
    got_names = [p.name for p in got]
    if name == 'SMPL.forward':                       # (*args, **kwargs) in the reference: explicit keywords here
        assert got_names[0] == 'self' and got[-1].kind is inspect.Parameter.VAR_KEYWORD
        return
    if name in MODULE_LEVEL:                         # state the reference keeps on `self` is passed explicitly
        assert [p for p in got_names if p in ref_params] == ref_params, (got_names, ref_params)
        return
    if name in CTOR_FREE:
        assert got_names[0] == 'self'
        common = [p for p in ref_params if p in got_names]
        assert common == [p for p in got_names if p in ref_params], (got_names, ref_params)   # same relative order
        return
    n = len(ref_params)
    assert got_names[:n] == ref_params, (key, got_names, ref_params)
    for p, d in zip(got[:n], ref_defaults):
        have = _default_src(p.default)
        if d is None:
            assert have is None, (key, p.name, have)
        else:
            assert have is not None, (key, p.name)
            try:                                     # numerically equal defaults (1e-2 vs 0.01, 5000 vs 5000.0)
                assert float(eval(d)) == float(eval(have)), (key, p.name, d, have)
            except (TypeError, ValueError, NameError, SyntaxError):
                assert d.replace('"', "'") == have.replace('"', "'"), (key, p.name, d, have)


def test_reference_module_paths_alias_to_mirrors():
    """The binding of INTEGRATION.md: after aliasing, the reference's own import statements resolve to the
    sm_100a path."""
    import sys
    names = ('utils.contact', 'utils.segmentation', 'utils.geometry', 'models.smpl', 'smplify.prior',
             'smplify.losses', 'smplify.smplifydc', 'train.loss', 'train.fits_dict')
    saved = {k: sys.modules.get(k) for k in ['tuch'] + ['tuch.' + n for n in names]}
    try:
        for n in names:
            sys.modules['tuch.' + n] = importlib.import_module('tuch_b200.' + n)
        from tuch.smplify.smplifydc import SMPLifyDC          # train.py:30, demo_smplify_dc.py:32
        from tuch.models.smpl import SMPL                     # train.py:34
        from tuch.train.loss import RegressorLoss             # train.py:33
        from tuch.utils.contact import batch_pairwise_dist, winding_numbers      # losses.py:20,23
        from tuch.utils.segmentation import BatchBodySegment  # demo_smplify_dc.py:38
        import tuch_b200.smplify.smplifydc as m
        assert SMPLifyDC is m.SMPLifyDC and SMPL.__module__ == 'tuch_b200.models.smpl'
        assert RegressorLoss.__module__ == 'tuch_b200.train.loss' and callable(batch_pairwise_dist)
        assert callable(winding_numbers) and BatchBodySegment.__module__ == 'tuch_b200.utils.segmentation'
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
