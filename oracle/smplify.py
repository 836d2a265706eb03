"""TEST INFRASTRUCTURE -- restatement of the two-stage SMPLify-DC loop,
tuch/smplify/smplifydc.py:68-236 (contact branch and SPIN branch) and :238-276, on torch CPU
with torch.optim.Adam (smplifydc.py:117,150,197: lr=step_size, betas=(0.9,0.999), eps=1e-8)."""
import numpy as np
import torch

from . import lbs as olbs
from . import losses as ol


def smplify_dc(model, prior, geomask_np, init_pose, init_betas, init_cam_t, center, keypoints_2d,
               ign_joints, num_iters=10, step_size=1e-2, focal_length=5000.0, euclthres=0.0,
               use_contact=True, cdict=None, gt_contact=None, ignore_idxs=None,
               has_discrete_contact=None, has_gt_keypoints=None, contact_loss_weight=1.0,
               segments=None, trace=None):
    """Returns (vertices, joints, pose, betas, cam_t, reprojection_loss[B,49], losses per iter)."""
    faces = model['faces'].numpy()
    cam_t = init_cam_t.clone()
    j2d = keypoints_2d[:, :, :2]
    conf = keypoints_2d[:, :, -1].clone()
    body_pose = init_pose[:, 3:].detach().clone()
    orient = init_pose[:, :3].detach().clone()
    betas = init_betas.detach().clone()

    def fwd():
        return olbs.smpl_forward(model, betas, body_pose, orient)

    # ---- stage 1: camera (+ shape | + global orientation)      smplifydc.py:100-134
    cam_t.requires_grad_(True)
    if use_contact:
        betas.requires_grad_(True)
        params = [betas, cam_t]
    else:
        orient.requires_grad_(True)
        params = [orient, cam_t]
    opt = torch.optim.Adam(params, lr=step_size, betas=(0.9, 0.999))
    hist = []
    for _ in range(num_iters):
        _, joints, _ = fwd()
        loss = ol.camera_fitting_loss(joints, betas, cam_t, init_cam_t, center, j2d, conf,
                                      focal_length=focal_length,
                                      shape_prior_weight=1.0 if use_contact else 0.0)
        opt.zero_grad()
        loss.backward()
        opt.step()
        hist.append(float(loss))

    # ---- stage 2                                                smplifydc.py:136-210
    conf[:, ign_joints] = 0.0
    if use_contact:
        cam_t.requires_grad_(False)
        betas.requires_grad_(False)
        body_pose.requires_grad_(True)
        orient.requires_grad_(True)
        opt = torch.optim.Adam([body_pose, orient], lr=step_size)
    else:
        body_pose.requires_grad_(True)
        betas.requires_grad_(True)
        orient.requires_grad_(True)
        cam_t.requires_grad_(False)
        opt = torch.optim.Adam([body_pose, betas, orient], lr=step_size, betas=(0.9, 0.999))
    for it in range(num_iters):
        verts, joints, _ = fwd()
        if use_contact:
            loss = ol.contact_fitting_loss(body_pose, betas, joints, geomask_np, euclthres, cam_t, center,
                                           j2d, conf, prior, cdict, gt_contact, ignore_idxs,
                                           has_discrete_contact, verts, faces, focal_length=focal_length,
                                           contact_loss_weight=contact_loss_weight, segments=segments)
        else:
            loss = ol.body_fitting_loss(body_pose, betas, joints, cam_t, center, j2d, conf, prior,
                                        focal_length=focal_length)
        opt.zero_grad()
        loss.backward()
        if trace is not None:
            trace.append(dict(loss=float(loss), body_pose=body_pose.detach().clone(),
                              orient=orient.detach().clone(),
                              g_body_pose=body_pose.grad.detach().clone(),
                              g_orient=orient.grad.detach().clone()))
        opt.step()
        hist.append(float(loss))

    # ---- final score                                            smplifydc.py:215-229
    with torch.no_grad():
        verts, joints, _ = fwd()
        if has_gt_keypoints is not None:
            conf[has_gt_keypoints, :25] = 0
        rep = ol.body_fitting_loss(body_pose, betas, joints, cam_t, center, j2d, conf, prior,
                                   focal_length=focal_length, output='reprojection')
    pose = torch.cat([orient, body_pose], dim=-1).detach()
    return verts.detach(), joints.detach(), pose, betas.detach(), cam_t.detach(), rep, hist
