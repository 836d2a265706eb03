import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tuch_b200 import synthetic as syn
from tuch_b200.models.smpl import SMPL
from oracle import lbs as olbs
torch.set_printoptions(precision=4, linewidth=200)
m = syn.make_body_model(10, 12)
dev = torch.device('cuda:0')
B = 3
pose = syn.fold_arms_pose(B, seed=1); pose[0] = 0
betas = np.random.default_rng(1).normal(0, .7, size=(B, 10)).astype(np.float32)
V = len(m['v_template'])
for mode in ('verts', 'joints'):
    tm = olbs.to_torch_model(m, torch.float64)
    p64 = torch.tensor(pose, dtype=torch.float64, requires_grad=True)
    v64, j64, _ = olbs.smpl_forward(tm, torch.tensor(betas, dtype=torch.float64), p64[:, 3:], p64[:, :3])
    (v64.sum() if mode == 'verts' else j64.sum()).backward()
    smpl = SMPL(model_arrays=m, batch_size=B).to(dev)
    p = torch.tensor(pose, device=dev, requires_grad=True)
    o = smpl(betas=torch.tensor(betas, device=dev), body_pose=p[:, 3:], global_orient=p[:, :3])
    (o.vertices.sum() if mode == 'verts' else o.joints.sum()).backward()
    d = (p.grad.cpu().double() - p64.grad).view(B, 24, 3)
    print(mode, 'max err per joint\n', d.abs().amax(-1))
    print('ref mag per joint\n', p64.grad.view(B, 24, 3).abs().amax(-1))
