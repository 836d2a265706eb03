"""GPU parity of the fused SMPL LBS forward/backward against the torch restatement of
smplx==0.1.13 (oracle.lbs; fp64 autograd is the gradient reference).  Tolerances: vertices and
joints 2e-6 m absolute (fp32 round-off of a metre-scale chain), gradients 2e-4 relative."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(batch, seed):
    from tuch_b200 import synthetic as syn
    rng = np.random.default_rng(seed)
    pose = syn.fold_arms_pose(batch, seed=seed)
    pose[0] = 0.0                                  # exact rest pose exercises the theta -> 0 branch
    betas = rng.normal(0, 0.7, size=(batch, 10)).astype(np.float32)
    return pose, betas


def _run(assets, batch, seed, rotmat):
    from oracle import lbs as olbs
    from tuch_b200.models.smpl import SMPL
    dev = torch.device('cuda:0')
    m = assets['model']
    pose, betas = _inputs(batch, seed)
    V = len(m['v_template'])
    rng = np.random.default_rng(seed + 1)
    wv = rng.normal(size=(batch, V, 3))
    wj = rng.normal(size=(batch, 49, 3))

    tm64 = olbs.to_torch_model(m, torch.float64)
    b64 = torch.tensor(betas, dtype=torch.float64, requires_grad=True)
    if rotmat:
        R = olbs.rodrigues(torch.tensor(pose, dtype=torch.float64).reshape(-1, 3)).view(batch, 24, 3, 3)
        go64 = R[:, :1].clone().requires_grad_(True)
        bp64 = R[:, 1:].clone().requires_grad_(True)
    else:
        go64 = torch.tensor(pose[:, :3], dtype=torch.float64, requires_grad=True)
        bp64 = torch.tensor(pose[:, 3:], dtype=torch.float64, requires_grad=True)
    v64, j64, _ = olbs.smpl_forward(tm64, b64, bp64, go64, pose2rot=not rotmat)
    ((v64 * torch.tensor(wv)).sum() + (j64 * torch.tensor(wj)).sum()).backward()

    smpl = SMPL(model_arrays=m, batch_size=batch).to(dev)
    bt = torch.tensor(betas, device=dev, requires_grad=True)
    go = go64.detach().float().to(dev).requires_grad_(True)
    bp = bp64.detach().float().to(dev).requires_grad_(True)
    out = smpl(betas=bt, body_pose=bp, global_orient=go, pose2rot=not rotmat, return_full_pose=True)
    assert out.vertices.shape == (batch, V, 3) and out.joints.shape == (batch, 49, 3)
    assert (out.vertices.detach().cpu().double() - v64.detach()).abs().max() < 2e-6
    assert (out.joints.detach().cpu().double() - j64.detach()).abs().max() < 2e-6
    ((out.vertices * torch.tensor(wv, device=dev, dtype=torch.float32)).sum()
     + (out.joints * torch.tensor(wj, device=dev, dtype=torch.float32)).sum()).backward()
    for got, ref, name in ((bt.grad, b64.grad, 'betas'), (bp.grad, bp64.grad, 'body_pose'),
                           (go.grad, go64.grad, 'global_orient')):
        err = (got.cpu().double() - ref).abs().max() / ref.abs().max()
        assert err < 2e-4, (name, float(err))
    return out


def test_lbs_small_axis_angle(small_assets):
    _run(small_assets, batch=3, seed=1, rotmat=False)
    _run(small_assets, batch=11, seed=2, rotmat=False)


def test_lbs_small_rotmat(small_assets):
    _run(small_assets, batch=5, seed=3, rotmat=True)


def test_lbs_full_size(full_assets):
    _run(full_assets, batch=9, seed=4, rotmat=False)
    _run(full_assets, batch=2, seed=5, rotmat=True)


def test_lbs_partial_gradients_and_defaults(small_assets):
    """vertices-only / joints-only cotangents, frozen inputs, default (omitted) arguments."""
    from oracle import lbs as olbs
    from tuch_b200.models.smpl import SMPL
    dev = torch.device('cuda:0')
    m = small_assets['model']
    smpl = SMPL(model_arrays=m, batch_size=2).to(dev)
    out = smpl()                                                     # rest pose, zero betas
    assert (out.vertices[0].cpu() - torch.tensor(m['v_template'])).abs().max() < 1e-6
    pose, betas = _inputs(2, 7)
    tm = olbs.to_torch_model(m, torch.float64)
    bp64 = torch.tensor(pose[:, 3:], dtype=torch.float64, requires_grad=True)
    v64, j64, _ = olbs.smpl_forward(tm, torch.tensor(betas, dtype=torch.float64), bp64,
                                    torch.tensor(pose[:, :3], dtype=torch.float64))
    j64.square().sum().backward()
    bp = torch.tensor(pose[:, 3:], device=dev, requires_grad=True)
    o = smpl(betas=torch.tensor(betas, device=dev), body_pose=bp, global_orient=torch.tensor(pose[:, :3], device=dev))
    o.joints.square().sum().backward()
    assert (bp.grad.cpu().double() - bp64.grad).abs().max() / bp64.grad.abs().max() < 2e-4
    assert smpl.faces.shape == m['faces'].shape and smpl.get_num_verts() == len(m['v_template'])
