"""GPU parity of the TUCH train-step mirror (tuch_b200/train/train_module.py: forward_train_step,
get_verts_in_contact) against the CPU restatement of tuch/train/train_module.py:93-336 in oracle/train_step.py,
on a mixed dsc / mtp batch with SMPLify-DC in the loop and a stand-in image regressor."""
from collections import namedtuple

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
Opt = namedtuple('Opt', ['batch_size', 'img_res', 'run_smplify', 'use_contact_in_the_loop',
                         'contact_in_the_loop_loss_weight', 'smplify_threshold', 'contact_loss_weight',
                         'openpose_train_weight', 'gt_train_weight', 'shape_loss_weight', 'keypoint_loss_weight',
                         'pose_loss_weight', 'beta_loss_weight'])
B, ITERS, GEO_FIT, GEO_CRIT, EUCL = 4, 3, 0.3, 0.3, 0.02


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def options(run_smplify=True):
    return Opt(B, 224, run_smplify, True, 2000.0, 10.5, 1.0, 0.0, 1.0, 0.5, 5.0, 1.0, 0.001)


def make_inputs(a, seed=0):
    from oracle import lbs as olbs
    from tuch_b200 import synthetic as syn
    tm = olbs.to_torch_model(a['model'])
    jf = lambda p, b: olbs.smpl_forward(tm, torch.tensor(b), torch.tensor(p[:, 3:]), torch.tensor(p[:, :3]))[1].numpy()
    batch, store = syn.make_train_batch(a['model'], a['regions'], B, seed=seed, joints_fn=jf, img_hw=16)
    return tm, batch, store


def run_oracle(a, tm, batch, store, o, net):
    from oracle import losses as ol, segments as oseg, train_step as ots
    from tuch_b200 import synthetic as syn
    from tuch_b200.train.fits_dict import SMPL_POSE_FLIP_PERM
    cb = {k: (v if k == 'dataset_name' else torch.tensor(v)) for k, v in batch.items()}
    segs = oseg.build_segments(a['segs'], a['model']['faces'])
    st = torch.tensor(store)
    out = ots.forward_train_step(
        o, cb, st, torch.tensor(SMPL_POSE_FLIP_PERM), tm, ol.GMMPrior(a['gmm']), torch.tensor(a['geo']), net,
        a['regions'], segs, [syn.JOINT_IDS[n] for n in syn.IGN_JOINTS], syn.FOCAL_LENGTH,
        dict(step_size=1e-2, num_iters=ITERS, geothres=GEO_FIT, euclthres=EUCL),
        dict(geothres=GEO_CRIT, euclthres=EUCL, hd_regressor=a['hd_reg'], hd_face_idx=a['hd_fidx'], use_hd=True))
    return out, st


def build_mirror(a, store, o, net):
    from tuch_b200 import synthetic as syn
    from tuch_b200.models.smpl import SMPL
    from tuch_b200.smplify.prior import MaxMixturePrior
    from tuch_b200.smplify.smplifydc import SMPLifyDC
    from tuch_b200.train.fits_dict import FitsDict
    from tuch_b200.train.loss import RegressorLoss
    from tuch_b200.train.train_module import TUCH
    from tuch_b200.utils.segmentation import BatchBodySegment
    m = a['model']
    dev = torch.device(DEV)
    faces = torch.tensor(m['faces'], device=DEV)
    face_tensor = faces[None].repeat(B, 1, 1)
    geod = torch.tensor(a['geo'], device=DEV)
    smpl = SMPL(model_arrays=m, batch_size=B).to(DEV)
    segments = BatchBodySegment(list(a['segs'].keys()), faces, segment_data=a['segs'])
    crit = RegressorLoss(o, DEV, len(m['v_template']), face_tensor, geod, geothres=GEO_CRIT, euclthres=EUCL,
                         face_tensor=face_tensor, use_hd=True, hd_regressor=a['hd_reg'], hd_faces=a['hd_fidx'],
                         segments=segments)
    smplify = SMPLifyDC(step_size=1e-2, batch_size=B, num_iters=ITERS, focal_length=syn.FOCAL_LENGTH, geodistssmpl=geod,
                        geothres=GEO_FIT, euclthres=EUCL, device=dev, smpl=SMPL(model_arrays=m, batch_size=B).to(DEV),
                        pose_prior=MaxMixturePrior(gmm=a['gmm'], num_gaussians=8).to(DEV),
                        ign_joints=[syn.JOINT_IDS[n] for n in syn.IGN_JOINTS])
    fits = FitsDict(device=dev, dataset_sizes={'dsc': len(store)})
    fits.fits_dict['dsc'] = torch.tensor(store)
    return TUCH(o, dev, None, smpl, None, net, smplify, crit, geod, fits_dict=fits, contactlists=a['regions'],
                focal_length=syn.FOCAL_LENGTH, geothres=GEO_FIT, euclthres=EUCL)


@pytest.mark.parametrize('run_smplify', [True, False])
def test_forward_train_step_against_oracle(small_assets, run_smplify):
    import copy
    from tuch_b200 import synthetic as syn
    a = small_assets
    o = options(run_smplify)
    tm, batch, store = make_inputs(a)
    net_cpu = syn.make_stand_in_regressor()
    net_gpu = copy.deepcopy(net_cpu).to(DEV)

    (loss_o, losses_o, out_o), store_o = run_oracle(a, tm, batch, store.copy(), o, net_cpu)
    loss_o.backward()

    tuch = build_mirror(a, store.copy(), o, net_gpu)
    gb = {k: (v if k == 'dataset_name' else torch.tensor(v, device=DEV)) for k, v in batch.items()}
    loss, losses, out = tuch.forward_train_step(gb)
    loss.backward()

    assert set(losses) == {'loss', 'loss_shape', 'loss_keypoints', 'loss_keypoints_3d', 'loss_regr_pose',
                           'loss_regr_betas', 'loss_cam', 'loss_contact'}
    assert set(out) == {'pred_vertices', 'spin_vertices', 'opt_vertices', 'pred_cam_t', 'spin_cam_t', 'opt_cam_t',
                        'smplifyoptiverts', 'gt_contact_l3', 'has_contact_pc', 'has_contact', 'valid_kpts_anno',
                        'gt_keypoints'}
    assert float(losses_o['loss_contact']) > 0                              # the predictions do self-intersect
    for k in losses:
        assert rel(losses[k], losses_o[k]) < 5e-4, (k, float(losses[k]), float(losses_o[k]))
    assert torch.equal(out['valid_kpts_anno'].cpu(), out_o['valid_kpts_anno'])
    for k in ('pred_vertices', 'opt_vertices', 'pred_cam_t', 'opt_cam_t', 'gt_keypoints'):
        assert rel(out[k], out_o[k]) < 5e-4, k
    # the fits store received the same rows (FitsDict.__setitem__ un-flips and rotates back)
    new_store = tuch.fits_dict.fits_dict['dsc']
    changed_o = (store_o != torch.tensor(store)).any(dim=1)
    changed = (new_store != torch.tensor(store)).any(dim=1)
    assert torch.equal(changed, changed_o)
    assert bool(changed.any()) == run_smplify
    assert rel(new_store, store_o) < 1e-3
    assert (len(out['smplifyoptiverts']) == ITERS) if run_smplify else (out['smplifyoptiverts'] is None)
    # gradient that reaches the regressor
    assert rel(net_gpu.fc.weight.grad, net_cpu.fc.weight.grad) < 2e-3
    assert rel(net_gpu.fc.bias.grad, net_cpu.fc.bias.grad) < 2e-3


@pytest.mark.parametrize('tag,run_smplify', [('fit', True), ('nofit', False)])
def test_forward_train_step_matches_reference_golden(small_assets, tag, run_smplify):
    """The train-step mirror against what the reference's OWN TUCH.forward_train_step (tuch/train/train_module.py:
    112-336, with its own SMPLifyDC, RegressorLoss, FitsDict and estimate_translation) produced on the same batch:
    tests/golden/train_step.npz, recorded by tests/golden/make_golden_train.py."""
    from conftest import golden
    from tuch_b200 import synthetic as syn
    g = golden('train_step.npz')
    a = small_assets
    o = options(run_smplify)
    tm, batch, store = make_inputs(a)
    net = syn.make_stand_in_regressor().to(DEV)
    tuch = build_mirror(a, store.copy(), o, net)
    gb = {k: (v if k == 'dataset_name' else torch.tensor(v, device=DEV)) for k, v in batch.items()}
    loss, losses, out = tuch.forward_train_step(gb)
    loss.backward()
    for k, v in losses.items():
        ref = g['%s/losses/%s' % (tag, k)]
        assert rel(v.reshape(-1), ref) < 5e-4, (k, float(v.reshape(-1)[0]), ref)
    assert np.array_equal(out['valid_kpts_anno'].cpu().numpy(), g[tag + '/out/valid_kpts_anno'])
    for k in ('pred_vertices', 'opt_vertices', 'pred_cam_t', 'opt_cam_t', 'gt_keypoints'):
        assert rel(out[k], g['%s/out/%s' % (tag, k)]) < 5e-4, k
    new_store = tuch.fits_dict.fits_dict['dsc'].cpu().numpy()
    assert np.array_equal((new_store != store).any(axis=1), (g[tag + '/store'] != g['store']).any(axis=1))
    assert rel(new_store, g[tag + '/store']) < 1e-3
    assert rel(net.fc.weight.grad, g[tag + '/g_weight']) < 2e-3
    assert rel(net.fc.bias.grad, g[tag + '/g_bias']) < 2e-3
    n = 0 if out['smplifyoptiverts'] is None else len(out['smplifyoptiverts'])
    assert n == int(g[tag + '/n_optiverts'])


def test_get_verts_in_contact_against_oracle(small_assets):
    """train_module.py:93-110 per body on CPU: rows with a partner closer than euclthres among the vertices
    at least geothres away along the surface, and the first closest such partner."""
    from oracle import clib
    a = small_assets
    o = options(False)
    tm, batch, store = make_inputs(a)
    tuch = build_mirror(a, store, o, None)
    from tuch_b200 import synthetic as syn
    from oracle import lbs as olbs
    pose = torch.tensor(syn.fold_arms_pose(3, seed=4, fold=1.2))
    verts = olbs.smpl_forward(tm, torch.zeros(3, 10), pose[:, 3:], pose[:, :3])[0]
    got = tuch.get_verts_in_contact(verts.to(DEV))
    mask = a['geo'] >= GEO_FIT
    total = 0
    for b in range(3):
        am, mn = clib.masked_nearest(verts[b].numpy(), mask)
        rows = np.where(mn < EUCL ** 2)[0]
        assert np.array_equal(got[b][0].cpu().numpy(), rows)
        assert np.array_equal(got[b][1].cpu().numpy(), am[rows])
        total += len(rows)
    assert total > 0
