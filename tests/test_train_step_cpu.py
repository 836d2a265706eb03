"""Pins oracle/train_step.py (the CPU restatement of TUCH.forward_train_step) against outputs of the reference's OWN
tuch/train/train_module.py:112-336 recorded by tests/golden/make_golden_train.py: losses, the regressor's gradient,
the supervision flags, the optimised bodies and the rows written back to the fits store."""
import numpy as np
import pytest
import torch

from conftest import golden
from test_train_step_gpu import make_inputs, options, rel, run_oracle


@pytest.mark.parametrize('tag,run_smplify', [('fit', True), ('nofit', False)])
def test_oracle_train_step_matches_reference(small_assets, tag, run_smplify):
    from tuch_b200 import synthetic as syn
    g = golden('train_step.npz')
    a = small_assets
    tm, batch, store = make_inputs(a)
    for k, v in batch.items():                                   # the golden was recorded on these very inputs
        if k != 'dataset_name':
            assert np.array_equal(np.asarray(v), g['batch/' + k]), k
    net = syn.make_stand_in_regressor()
    (loss, losses, out), new_store = run_oracle(a, tm, batch, store.copy(), options(run_smplify), net)
    loss.backward()
    for k, v in losses.items():
        ref = g['%s/losses/%s' % (tag, k)]
        assert rel(v.reshape(-1), ref) < 1e-4, (k, float(v.reshape(-1)[0]), ref)
    assert np.array_equal(out['valid_kpts_anno'].numpy(), g[tag + '/out/valid_kpts_anno'])
    for k in ('pred_vertices', 'opt_vertices', 'pred_cam_t', 'opt_cam_t', 'gt_keypoints'):
        assert rel(out[k], g['%s/out/%s' % (tag, k)]) < 1e-4, k
    changed = (new_store != torch.tensor(store)).any(dim=1).numpy()
    assert np.array_equal(changed, (g[tag + '/store'] != g['store']).any(axis=1))
    assert rel(new_store, g[tag + '/store']) < 1e-4
    assert rel(net.fc.weight.grad, g[tag + '/g_weight']) < 1e-3
    assert rel(net.fc.bias.grad, g[tag + '/g_bias']) < 1e-3
    assert int(g[tag + '/n_optiverts']) == (3 if run_smplify else 0)       # one vertex set per stage-2 iteration
