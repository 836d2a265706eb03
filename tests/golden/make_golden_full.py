"""BASELINE config 1 at the real size: batch=1 self-contact push/pull loss of one SMPL-sized body
(V=6890, F=13776, lattice body) computed by EXECUTING the reference's own tuch/smplify/losses.py
(contact_fitting_loss, CPU-patched as in make_golden.py), tuch/utils/contact.py and, for two more bodies,
tuch/train/loss.py (RegressorLoss.contact_loss with a 20k-point HD model -> regressor_full_size.npz).  Needs ~8 GB of RAM
for the reference's [Q, F, 3, 3] solid-angle tensor.  Run in the build container:
    python tests/golden/make_golden_full.py   ->   tests/golden/contact_full_size.npz
"""
import functools
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, '/root/reference')

from tuch_b200 import synthetic as syn      # noqa: E402
from oracle import lbs as olbs              # noqa: E402
import make_golden as mg                    # noqa: E402


N_HD = 20000          # HD points of the regressor-loss golden (the paper's HD model has ~20 k)


def main():
    model = syn.make_lattice_body_model(seed=0)
    geo = syn.make_geodesics(model['v_template'], model['faces'], cache_dir='/tmp/tuch_b200_cache')
    regions = syn.make_regions(model)
    segs = syn.make_segments(model)
    gmm = syn.make_gmm()
    hd_reg, hd_fidx = syn.make_hd_regressor(model, n_hd=N_HD)
    cwd = os.getcwd()
    work = tempfile.mkdtemp(prefix='tuch_golden_full_')
    tmodel = mg.install_stubs(work, model, segs, hd_reg, hd_fidx, gmm)
    import tuch.utils.contact as rc
    import tuch.smplify.losses as rl
    import tuch.smplify.prior as rp
    import tuch.utils.segmentation as rseg
    from configs import config as rconfig
    cpu_pd = functools.partial(rc.batch_pairwise_dist, use_cuda=False)
    rl.batch_pairwise_dist = cpu_pd
    B = 1
    inp = syn.make_smplify_inputs(
        model, regions, B, seed=77,
        joints_fn=lambda p, b: olbs.smpl_forward(tmodel, torch.tensor(b), torch.tensor(p[:, 3:]), torch.tensor(p[:, :3]))[1].numpy())
    pose = torch.tensor(inp['init_pose'])
    betas = torch.tensor(inp['init_betas'])
    bp = pose[:, 3:].clone().requires_grad_(True)
    go = pose[:, :3].clone().requires_grad_(True)
    verts, joints, _ = olbs.smpl_forward(tmodel, betas, bp, go)
    verts.retain_grad()
    joints.retain_grad()
    faces = torch.tensor(model['faces'])
    face_tensor = faces[None]
    geothres = 0.3
    geomask = torch.tensor(geo) > geothres
    # the reference's segment class over the stubbed trimesh / segm_utils (segmentation.py:102-124)
    os.makedirs(rconfig.SEGMENT_DIR, exist_ok=True)
    segments = rseg.BatchBodySegment(list(segs.keys()), faces)
    prior = rp.MaxMixturePrior(prior_folder=os.path.dirname(rconfig.PRIOR_FOLDER + '/x'), num_gaussians=8, dtype=torch.float32)
    loss = rl.contact_fitting_loss(
        bp, go, bp.detach(), go.detach(), betas, joints, geomask, 0.02,
        torch.tensor(inp['init_cam_t']), torch.tensor(inp['camera_center']),
        torch.tensor(inp['keypoints_2d'][:, :, :2]), torch.tensor(inp['keypoints_2d'][:, :, 2]),
        prior, cdict=regions, gt_contact=[torch.tensor(inp['gt_contact']), None],
        ignore_idxs=torch.tensor(inp['ignore_idxs']), has_discrete_contact=torch.tensor(inp['has_discrete_contact']),
        verts=verts, face_tensor=face_tensor, device='cpu', focal_length=5000.0,
        contact_loss_weight=2000.0, segments=segments)
    loss.backward()
    with torch.no_grad():
        vv = verts.detach()
        wn = rc.winding_numbers(vv, vv[0][faces][None]).squeeze()
        P = cpu_pd(vv, vv, squared=True)
        P[:, ~geomask] = float('inf')
        am = torch.argmin(P, axis=1)[0]
        mn = P.min(1)[0][0]
    # ---- RegressorLoss.contact_loss with the HD path (tuch/train/loss.py:240-317) on two bodies
    import tuch.train.loss as rtl
    from collections import namedtuple
    rtl.batch_pairwise_dist = cpu_pd
    Opt = namedtuple('Opt', ['contact_loss_weight'])
    crit = rtl.RegressorLoss(Opt(1.0), torch.device('cpu'), len(model['v_template']), face_tensor.repeat(2, 1, 1),
                             torch.tensor(geo), geothres=geothres, euclthres=0.02,
                             face_tensor=face_tensor.repeat(2, 1, 1), use_hd=True)
    pose2 = torch.tensor(syn.fold_arms_pose(2, seed=78))
    verts2 = olbs.smpl_forward(tmodel, torch.zeros(2, 10), pose2[:, 3:], pose2[:, :3])[0].detach()
    pv = verts2.clone().requires_grad_(True)
    rval = crit.contact_loss(pv, torch.tensor([True, True]))
    rval.backward()
    os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, 'regressor_full_size.npz'), verts=verts2.numpy(), loss=rval.item(),
                        g_verts=pv.grad.numpy().astype(np.float32), n_hd=N_HD, geothres=geothres)
    print('regressor loss', rval.item(), 'file KB', os.path.getsize(os.path.join(HERE, 'regressor_full_size.npz')) // 1024)
    np.savez_compressed(os.path.join(HERE, 'contact_full_size.npz'),
                        init_pose=inp['init_pose'], init_betas=inp['init_betas'], init_cam_t=inp['init_cam_t'],
                        camera_center=inp['camera_center'], keypoints_2d=inp['keypoints_2d'],
                        gt_contact=inp['gt_contact'], has_discrete_contact=inp['has_discrete_contact'],
                        ignore_idxs=inp['ignore_idxs'], geothres=geothres,
                        verts=verts.detach().numpy(), joints=joints.detach().numpy(), loss=loss.item(),
                        g_verts=verts.grad.numpy().astype(np.float32), g_joints=joints.grad.numpy(),
                        g_body_pose=bp.grad.numpy(), g_orient=go.grad.numpy(),
                        winding=wn.numpy().astype(np.float32), argmin=am.numpy().astype(np.int32),
                        min_sq=mn.numpy().astype(np.float32), v_sum=float(model['v_template'].astype(np.float64).sum()))
    print('loss', loss.item(), 'interior', int((wn > 0.99).sum()), 'file KB',
          os.path.getsize(os.path.join(HERE, 'contact_full_size.npz')) // 1024)


if __name__ == '__main__':
    main()
