"""TEST INFRASTRUCTURE -- restatement of TUCH.forward_train_step, tuch/train/train_module.py:112-336, and
of the SPIN terms of RegressorLoss.forward, tuch/train/loss.py:94-238, on torch CPU, composed from the
other oracle modules (LBS, SMPLify-DC loop, regressor contact loss, pose bookkeeping).

This follows the reference's statements one by one; the image regressor is whatever callable the test passes.
PINNED: tests/golden/make_golden_train.py executes the reference's own forward_train_step (stub modules for the
absent smplx / trimesh / data tree, torchgeometry's two conversions from oracle/pose.py) and
tests/test_train_step_cpu.py holds this restatement to the recorded losses, gradients, supervision flags,
optimised bodies and fits-store rows.
"""
import numpy as np
import torch

from . import lbs as olbs
from . import pose as opose
from . import regressor as oreg
from . import smplify as osm


def quat_rodrigues(theta):
    """tuch/utils/geometry.py:29-65 (the quaternion form RegressorLoss.smpl_losses uses)."""
    angle = torch.norm(theta + 1e-8, p=2, dim=1, keepdim=True)
    axis = theta / angle
    q = torch.cat([torch.cos(0.5 * angle), torch.sin(0.5 * angle) * axis], dim=1)
    q = q / q.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).view(-1, 3, 3)


def regressor_loss(o, pred_rotmat, pred_betas, opt_pose, opt_betas, pred_kp2d, gt_kp2d, pred_joints, gt_joints,
                   has_pose_3d, pred_vertices, opt_vertices, pred_camera, valid_pose, valid_shape, contact_fn):
    """loss.py:94-168; contact_fn(pred_vertices, valid) is loss.py:240-317."""
    mse = torch.nn.MSELoss(reduction='none')
    loss_contact = contact_fn(pred_vertices, valid_pose) if o.contact_loss_weight > 0 else torch.tensor(0)
    # :216-238
    sp, ss = valid_pose == 1, valid_shape == 1
    gt_rotmat = quat_rodrigues(opt_pose.view(-1, 3)).view(-1, 24, 3, 3)
    zero = torch.zeros(1)
    l_pose = torch.nn.functional.mse_loss(pred_rotmat[sp], gt_rotmat[sp]) if int(sp.sum()) > 0 else zero
    l_betas = torch.nn.functional.mse_loss(pred_betas[ss], opt_betas[ss]) if int(ss.sum()) > 0 else zero
    # :170-182
    conf = gt_kp2d[:, :, -1].unsqueeze(-1).clone()
    conf[:, :25] *= o.openpose_train_weight
    conf[:, 25:] *= o.gt_train_weight
    l_kp = (conf * mse(pred_kp2d, gt_kp2d[:, :, :-1])).mean(dim=(1, 2))[valid_pose].mean()
    # :184-203
    sel = has_pose_3d == 1
    if int(sel.sum()) > 0:
        pred = pred_joints[:, 25:, :][sel]
        c3 = gt_joints[:, :, -1].unsqueeze(-1)[sel]
        gt = gt_joints[:, :, :-1][sel]
        gt = gt - ((gt[:, 2, :] + gt[:, 3, :]) / 2)[:, None, :]
        pred = pred - ((pred[:, 2, :] + pred[:, 3, :]) / 2)[:, None, :]
        l_kp3 = (c3 * mse(pred, gt)).mean()
    else:
        l_kp3 = zero
    # :205-214
    l_shape = torch.nn.functional.l1_loss(pred_vertices[sp], opt_vertices[sp]) if int(sp.sum()) > 0 else zero
    l_cam = ((torch.exp(-pred_camera[:, 0] * 10)) ** 2).mean()
    total = o.shape_loss_weight * l_shape + o.keypoint_loss_weight * l_kp + o.keypoint_loss_weight * l_kp3 + \
        o.pose_loss_weight * l_pose + o.beta_loss_weight * l_betas + l_cam + o.contact_loss_weight * loss_contact
    return total, {'loss_shape': l_shape, 'loss_keypoints': l_kp, 'loss_keypoints_3d': l_kp3,
                   'loss_regr_pose': l_pose, 'loss_regr_betas': l_betas, 'loss_cam': l_cam,
                   'loss_contact': loss_contact}


def forward_train_step(o, batch, store, flip_perm, model, prior, geodist, regressor, cdict, segments, ign_joints,
                       focal_length, smplify_kw, criterion_kw):
    """batch: dict of CPU tensors with the keys train_module.py:120-141 reads; store: [N,82] CPU tensor (the
    single dataset's fits, updated in place like FitsDict.__setitem__); model: oracle.lbs torch model;
    smplify_kw: step_size, num_iters, geothres, euclthres; criterion_kw: geothres, euclthres, hd_regressor,
    hd_face_idx, use_hd.  Returns (loss, losses, output) with the reference's keys."""
    faces = model['faces'].numpy()
    B = batch['img'].shape[0]
    idx, rot, flipped = batch['sample_index'], batch['rot_angle'], batch['is_flipped']
    has_pose_3d = batch['has_pose_3d'].bool()
    has_dc = batch['has_disc_contact'].bool()
    has_kp = batch['has_gt_kpts'].bool()
    has_smpl_ = batch['has_smpl'].bool() | batch['has_pgt_smpl'].bool()
    gt_kp, gt_joints, gt_pose, gt_betas, gt_dc = (batch[k] for k in ('keypoints', 'pose_3d', 'pose', 'betas',
                                                                     'contact_vec'))
    with torch.no_grad():
        gt_verts, gt_mj, _ = olbs.smpl_forward(model, gt_betas, gt_pose[:, 3:], gt_pose[:, :3])
    kp_px = gt_kp.clone()
    kp_px[:, :, :-1] = 0.5 * o.img_res * (kp_px[:, :, :-1] + 1)

    # FitsDict.__getitem__ (fits_dict.py:59-71): rotate, then flip
    params = store[idx.long()].clone()
    opt_pose = opose.flip_pose(opose.rotate_pose(params[:, :72].clone(), rot.float()), flipped, flip_perm).float()
    opt_betas = params[:, 72:].clone()
    with torch.no_grad():
        opt_verts, opt_joints, _ = olbs.smpl_forward(model, opt_betas, opt_pose[:, 3:], opt_pose[:, :3])
    opt_verts, opt_joints = opt_verts.clone(), opt_joints.clone()
    opt_c3 = torch.from_numpy(oreg.contact_from_verts(opt_verts.numpy(), cdict))

    gt_cam_t = opose.estimate_translation(gt_mj, kp_px, focal_length=focal_length, img_size=o.img_res,
                                          has_2d_kp_anno=has_kp)
    opt_cam_t = opose.estimate_translation(opt_joints, kp_px, focal_length=focal_length, img_size=o.img_res,
                                           has_2d_kp_anno=has_kp)
    center = 0.5 * o.img_res * torch.ones(B, 2)

    # get_fitting_loss (smplifydc.py:238-276): zeroes the ignored confidences in kp_px itself
    kp_px[:, ign_joints, -1] = 0.
    conf = kp_px[:, :, -1].clone()
    conf[has_kp, :25] = 0
    from . import losses as ol
    opt_joint_loss = ol.body_fitting_loss(opt_pose[:, 3:], opt_betas, opt_joints, opt_cam_t, center, kp_px[:, :, :2],
                                          conf, prior, focal_length=focal_length, output='reprojection').mean(dim=-1)

    pred_rotmat, pred_betas, pred_camera = regressor(batch['img'])
    pred_vertices, pred_joints, _ = olbs.smpl_forward(model, pred_betas, pred_rotmat[:, 1:],
                                                      pred_rotmat[:, 0].unsqueeze(1), pose2rot=False)
    hom = torch.cat([pred_rotmat.detach().view(-1, 3, 3),
                     torch.tensor([0, 0, 1], dtype=torch.float32).view(1, 3, 1).expand(B * 24, -1, -1)], dim=-1)
    pred_pose = opose.rotation_matrix_to_angle_axis(hom).contiguous().view(B, -1)
    pred_pose[torch.isnan(pred_pose)] = 0.0
    pred_cam_t = torch.stack([pred_camera[:, 1], pred_camera[:, 2],
                              2 * focal_length / (o.img_res * pred_camera[:, 0] + 1e-9)], dim=-1)
    p = pred_joints + pred_cam_t.unsqueeze(1)
    pred_kp2d = (focal_length * p[:, :, :2] / p[:, :, 2:3]) / (o.img_res / 2.)

    optiverts = None
    if o.run_smplify:
        geomask = (geodist > smplify_kw['geothres']).numpy()
        nv, nj, npose, nbetas, ncam, nloss, _ = osm.smplify_dc(
            model, prior, geomask, pred_pose.detach(), pred_betas.detach(), pred_cam_t.detach(), center, kp_px,
            ign_joints, num_iters=smplify_kw['num_iters'], step_size=smplify_kw['step_size'],
            focal_length=focal_length, euclthres=smplify_kw['euclthres'], use_contact=o.use_contact_in_the_loop,
            cdict=cdict, gt_contact=gt_dc.numpy(), ignore_idxs=has_smpl_, has_discrete_contact=has_dc,
            has_gt_keypoints=has_kp, contact_loss_weight=o.contact_in_the_loop_loss_weight, segments=segments)
        nloss = nloss.mean(dim=-1)
        update = nloss <= opt_joint_loss
        new_c3 = torch.from_numpy(oreg.contact_from_verts(nv.numpy(), cdict))
        upd_c3 = ((gt_dc * new_c3) <= (gt_dc * opt_c3)).sum(1) > 0
        if o.use_contact_in_the_loop:
            update[has_dc] = (upd_c3 * update)[has_dc]
        opt_joint_loss[update] = nloss[update]
        opt_verts[update] = nv[update]
        opt_c3[update] = new_c3[update]
        opt_joints[update] = nj[update]
        opt_pose[update] = npose[update]
        opt_betas[update] = nbetas[update]
        opt_cam_t[update] = ncam[update]
        # FitsDict.__setitem__ (fits_dict.py:73-87): un-flip, then rotate back
        back = opose.rotate_pose(opose.flip_pose(opt_pose.clone(), flipped, flip_perm), -rot.float()).float()
        new_params = torch.cat([back, opt_betas], dim=-1)
        for n in range(B):
            if update[n]:
                store[int(idx[n])] = new_params[n]

    opt_cam_t[has_smpl_] = gt_cam_t[has_smpl_]
    opt_joints[has_smpl_] = gt_mj[has_smpl_]
    opt_pose[has_smpl_] = gt_pose[has_smpl_]
    opt_betas[has_smpl_] = gt_betas[has_smpl_]
    opt_verts[has_smpl_] = gt_verts[has_smpl_]
    valid_fit = opt_joint_loss < o.smplify_threshold
    valid = has_smpl_ | valid_fit

    geomask_c = (geodist > criterion_kw['geothres']).numpy()

    def contact_fn(pv, vf):
        return oreg.regressor_contact_loss(pv, vf.numpy(), faces, geomask_c, criterion_kw['euclthres'], segments,
                                           criterion_kw.get('hd_regressor'), criterion_kw.get('hd_face_idx'),
                                           use_hd=criterion_kw.get('use_hd', True))
    loss, ld = regressor_loss(o, pred_rotmat, pred_betas, opt_pose, opt_betas, pred_kp2d, gt_kp, pred_joints, gt_joints,
                              has_pose_3d, pred_vertices, opt_verts, pred_camera, valid, valid, contact_fn)
    losses = {'loss': loss.detach(), **{k: v.detach() for k, v in ld.items()}}
    output = {'pred_vertices': pred_vertices.detach(), 'opt_vertices': opt_verts, 'pred_cam_t': pred_cam_t.detach(),
              'opt_cam_t': opt_cam_t, 'valid_kpts_anno': valid, 'gt_keypoints': kp_px, 'opt_pose': opt_pose,
              'opt_betas': opt_betas, 'opt_joint_loss': opt_joint_loss}
    return loss, losses, output
