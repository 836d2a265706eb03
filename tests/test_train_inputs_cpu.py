"""CPU checks around the train-step mirror: the synthetic batch follows the schema the reference's dataset
emits (tuch/datasets/base_dataset.py:310-331 as read by tuch/train/train_module.py:120-141),
the stand-in regressor returns proper rotations, and the mirror refuses to run without a CUDA device."""
import numpy as np
import pytest
import torch


def test_train_batch_schema(small_assets):
    from tuch_b200 import synthetic as syn
    a = small_assets
    B = 6
    batch, store = syn.make_train_batch(a['model'], a['regions'], B, seed=3, img_hw=16)
    n_cls = len(a['regions']['classes'])
    shapes = {'img': (B, 3, 16, 16), 'keypoints': (B, 49, 3), 'pose_3d': (B, 24, 4), 'pose': (B, 72), 'betas': (B, 10),
              'contact_vec': (B, n_cls), 'has_smpl': (B,), 'has_pgt_smpl': (B,), 'has_disc_contact': (B,),
              'has_gt_kpts': (B,), 'has_pose_3d': (B,), 'is_flipped': (B,), 'rot_angle': (B,), 'sample_index': (B,)}
    for k, shp in shapes.items():
        assert batch[k].shape == shp, k
    assert len(batch['dataset_name']) == B and store.shape == (2 * B, 82)
    dsc = batch['has_disc_contact'].astype(bool)
    # "dsc" rows: contact labels and no SMPL ground truth; "mtp" rows: pseudo ground truth and no labels
    assert np.array_equal(dsc, ~batch['has_pgt_smpl'].astype(bool))
    assert (batch['contact_vec'][dsc].sum(1) >= 1).all() and (batch['contact_vec'][~dsc] == 0).all()
    assert (batch['pose'][dsc] == 0).all() and (np.abs(batch['pose'][~dsc]).sum(1) > 0).all()
    assert np.abs(batch['keypoints'][:, :, :2]).max() <= 1.5 and (batch['keypoints'][:, :, 2] > 0).all()
    assert len(set(batch['sample_index'].tolist())) == B and batch['sample_index'].max() < len(store)
    assert not batch['has_gt_kpts'].astype(bool)[~dsc].any()


def test_stand_in_regressor_outputs_rotations():
    from tuch_b200 import synthetic as syn
    net = syn.make_stand_in_regressor(seed=1)
    rot, betas, cam = net(torch.randn(5, 3, 32, 32, generator=torch.Generator().manual_seed(0)))
    assert rot.shape == (5, 24, 3, 3) and betas.shape == (5, 10) and cam.shape == (5, 3)
    eye = torch.eye(3).expand(5, 24, 3, 3)
    assert torch.allclose(rot @ rot.transpose(-1, -2), eye, atol=1e-5)
    assert torch.allclose(torch.det(rot), torch.ones(5, 24), atol=1e-5)
    assert (cam[:, 0] > 0.5).all()
    (rot.sum() + betas.sum() + cam.sum()).backward()
    assert net.fc.weight.grad is not None and torch.isfinite(net.fc.weight.grad).all()


def test_train_step_mirror_needs_cuda():
    from tuch_b200.ops import TuchError
    from tuch_b200.train.train_module import TUCH
    with pytest.raises(TuchError, match='no CPU fallback'):
        TUCH(None, torch.device('cpu'), None, None, None, None, None, None, None, fits_dict=object(),
             contactlists={'classes': [], 'csig': {}}, focal_length=5000.0)


def test_train_step_mirror_reports_missing_data_tree():
    """Without contactlists= / focal_length= the constructor looks for the reference's data tree and says so
    when it is absent (no silent defaults)."""
    from tuch_b200.ops import TuchError
    from tuch_b200.train.train_module import TUCH
    if not torch.cuda.is_available():
        with pytest.raises(TuchError):
            TUCH(None, torch.device('cuda'), None, None, None, None, None, None, None)
    else:
        with pytest.raises(TuchError, match='focal_length'):
            TUCH(None, torch.device('cuda'), None, None, None, None, None, None, None)


def test_oracle_train_step_runs_and_keeps_its_invariants(small_assets):
    """The CPU restatement of TUCH.forward_train_step (oracle/train_step.py) on its own: finite losses with the
    reference's keys, SMPLify-DC in the loop only ever rewrites rows of the fits store that belong to the batch,
    ground-truth rows end up with their ground-truth parameters, and without fitting the store stays untouched."""
    import test_train_step_gpu as T
    from tuch_b200 import synthetic as syn
    a = small_assets
    tm, batch, store = T.make_inputs(a)
    for run_smplify in (True, False):
        net = syn.make_stand_in_regressor()
        (loss, losses, out), st = T.run_oracle(a, tm, batch, store.copy(), T.options(run_smplify), net)
        assert set(losses) == {'loss', 'loss_shape', 'loss_keypoints', 'loss_keypoints_3d', 'loss_regr_pose',
                               'loss_regr_betas', 'loss_cam', 'loss_contact'}
        assert all(bool(torch.isfinite(v).all()) for v in losses.values())
        changed = np.where((st.numpy() != store).any(axis=1))[0]
        assert set(changed.tolist()) <= set(batch['sample_index'].tolist())
        assert (len(changed) > 0) == run_smplify
        gt = batch['has_pgt_smpl'].astype(bool)
        assert np.array_equal(out['opt_pose'].numpy()[gt], batch['pose'][gt])
        assert np.array_equal(out['opt_betas'].numpy()[gt], batch['betas'][gt])
        assert bool(out['valid_kpts_anno'][torch.tensor(gt)].all())
        loss.backward()
        assert bool(torch.isfinite(net.fc.weight.grad).all()) and float(net.fc.weight.grad.abs().max()) > 0
