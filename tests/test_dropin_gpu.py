"""Drop-in acceptance: the reference's OWN caller code -- tuch/train/train_module.py
(TUCH.forward_train_step, :112-336) and tuch/train/fits_dict.py, staged UNMODIFIED from /root/reference into
baseline/_ref by scripts/stage_reference.py -- runs on the B200 against the tuch_b200 modules aliased onto the
reference's import paths exactly as INTEGRATION.md prescribes, and reproduces what the same caller code produced
over the reference's own modules (tests/golden/train_step.npz).  Skipped when the staged files are absent."""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import ROOT, golden
from test_train_step_gpu import B, DEV, EUCL, GEO_CRIT, GEO_FIT, ITERS, make_inputs, options, rel

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, 'baseline', '_ref')
ALIASES = ('utils.contact', 'utils.segmentation', 'utils.geometry', 'models.smpl', 'smplify.prior', 'smplify.losses',
           'smplify.smplifydc', 'train.loss')


@pytest.fixture()
def reference_caller():
    """The staged reference train_module / fits_dict imported over the aliased tuch_b200 modules."""
    if not os.path.exists(os.path.join(REF, 'tuch', 'train', 'train_module.py')):
        pytest.skip('baseline/_ref is not staged (scripts/stage_reference.py needs /root/reference)')
    from oracle import pose as opose                       # torchgeometry's conversions for the CPU fits store
    from tuch_b200 import synthetic as syn
    from tuch_b200.train.fits_dict import SMPL_POSE_FLIP_PERM
    from tuch_b200.utils import geometry as tgeo
    names = ['tuch', 'tuch.train', 'tuch.train.train_module', 'tuch.train.fits_dict', 'torchgeometry', 'configs',
             'configs.config', 'data', 'data.essentials', 'data.essentials.constants'] + ['tuch.' + n for n in ALIASES]
    saved = {k: sys.modules.get(k) for k in names}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m
    sys.path.insert(0, REF)
    try:
        for k in ('tuch', 'tuch.train', 'configs', 'configs.config'):
            sys.modules.pop(k, None)
        for n in ALIASES:                                  # the binding of INTEGRATION.md
            sys.modules['tuch.' + n] = importlib.import_module('tuch_b200.' + n)

        def r2aa(rotmat):                                  # train_module.py:211 on the device, fits_dict on the host
            return tgeo.rotation_matrix_to_angle_axis(rotmat) if rotmat.is_cuda else opose.rotation_matrix_to_angle_axis(rotmat)
        mod('torchgeometry', rotation_matrix_to_angle_axis=r2aa,
            angle_axis_to_rotation_matrix=opose.angle_axis_to_rotation_matrix)
        mod('data')
        mod('data.essentials')
        c = mod('data.essentials.constants', FOCAL_LENGTH=syn.FOCAL_LENGTH, SMPL_POSE_FLIP_PERM=SMPL_POSE_FLIP_PERM)
        sys.modules['data.essentials'].constants = c
        rtm = importlib.import_module('tuch.train.train_module')
        rfd = importlib.import_module('tuch.train.fits_dict')
        assert os.path.realpath(rtm.__file__).startswith(os.path.realpath(REF))
        yield rtm, rfd
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.parametrize('tag,run_smplify', [('fit', True), ('nofit', False)])
def test_reference_train_step_runs_unchanged_over_the_aliased_modules(small_assets, reference_caller, tag, run_smplify):
    rtm, rfd = reference_caller
    from tuch.models.smpl import SMPL                       # the reference's import paths, resolved to tuch_b200
    from tuch.smplify.prior import MaxMixturePrior
    from tuch.smplify.smplifydc import SMPLifyDC
    from tuch.train.loss import RegressorLoss
    from tuch.utils.segmentation import BatchBodySegment
    from tuch_b200 import synthetic as syn
    assert SMPLifyDC.__module__ == 'tuch_b200.smplify.smplifydc' and rtm.TUCH.__module__ == 'tuch.train.train_module'
    g = golden('train_step.npz')
    a, m = small_assets, small_assets['model']
    o = options(run_smplify)
    tm, batch, store = make_inputs(a)
    dev = torch.device(DEV)
    faces = torch.tensor(m['faces'], device=DEV)
    face_tensor = faces[None].repeat(B, 1, 1)
    geod = torch.tensor(a['geo'], device=DEV)
    segments = BatchBodySegment(list(a['segs'].keys()), faces, segment_data=a['segs'])
    net = syn.make_stand_in_regressor().to(DEV)
    fits = rfd.FitsDict.__new__(rfd.FitsDict)               # the reference's own fits store (CPU tensors, cv2)
    fits.flipped_parts = torch.tensor(list(sys.modules['data.essentials.constants'].SMPL_POSE_FLIP_PERM), dtype=torch.int64)
    fits.fits_dict = {'dsc': torch.tensor(store.copy())}
    tuch = rtm.TUCH.__new__(rtm.TUCH)                       # __init__ reads the un-shipped DSC pickles / datasets
    tuch.options, tuch.device, tuch.focal_length = o, dev, syn.FOCAL_LENGTH
    tuch.fits_dict, tuch.modelspin, tuch.model = fits, net, net
    tuch.smpl = SMPL(model_arrays=m, batch_size=B).to(DEV)
    tuch.geodistssmpl = geod
    tuch.smplify = SMPLifyDC(step_size=1e-2, batch_size=B, num_iters=ITERS, focal_length=syn.FOCAL_LENGTH,
                             geodistssmpl=geod, geothres=GEO_FIT, euclthres=EUCL, device=dev,
                             smpl=SMPL(model_arrays=m, batch_size=B).to(DEV),
                             pose_prior=MaxMixturePrior(gmm=a['gmm'], num_gaussians=8).to(DEV),
                             ign_joints=[syn.JOINT_IDS[n] for n in syn.IGN_JOINTS])
    tuch.criterion_cospin = RegressorLoss(o, DEV, len(m['v_template']), face_tensor, geod, geothres=GEO_CRIT,
                                          euclthres=EUCL, face_tensor=face_tensor, use_hd=True,
                                          hd_regressor=a['hd_reg'], hd_faces=a['hd_fidx'], segments=segments)
    tuch.contactlists = a['regions']
    gb = {k: (v if k == 'dataset_name' else torch.tensor(v, device=DEV)) for k, v in batch.items()}
    loss, losses, out = tuch.forward_train_step(gb)          # the reference's statements, :112-336, unchanged
    loss.backward()
    for k, v in losses.items():
        ref = g['%s/losses/%s' % (tag, k)]
        assert rel(v.reshape(-1), ref) < 5e-4, (k, float(v.reshape(-1)[0]), ref)
    assert np.array_equal(out['valid_kpts_anno'].cpu().numpy(), g[tag + '/out/valid_kpts_anno'])
    for k in ('pred_vertices', 'opt_vertices', 'pred_cam_t', 'opt_cam_t', 'gt_keypoints'):
        assert rel(out[k], g['%s/out/%s' % (tag, k)]) < 5e-4, k
    new_store = fits.fits_dict['dsc'].numpy()
    assert np.array_equal((new_store != store).any(axis=1), (g[tag + '/store'] != g['store']).any(axis=1))
    assert rel(new_store, g[tag + '/store']) < 1e-3
    assert rel(net.fc.weight.grad, g[tag + '/g_weight']) < 2e-3
    n = 0 if out['smplifyoptiverts'] is None else len(out['smplifyoptiverts'])
    assert n == int(g[tag + '/n_optiverts'])
