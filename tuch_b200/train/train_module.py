"""Host-side mirror of tuch/train/train_module.py: the TUCH train step around the contact hot path.

* contact_from_verts (TUCH.contact_from_verts :69-91), which the authors flag as a training-loop
  bottleneck ("Speed up this function will speed up training loop!", :74).  The reference loops over
  every annotated region pair in Python and runs three K=3 bmm's plus a flat min per pair; here one
  launch (tuch_region_min, unmasked) computes every (body, pair) minimum.
* get_verts_in_contact (:93-110): the masked nearest-vertex kernel instead of a [V,V] matrix per body.
* TUCH.forward_train_step (:112-336): the same sequence of calls -- fits lookup, two SMPL forwards,
  region contact, camera estimates, the regressor, SMPLify-DC in the loop, the fits update and
  RegressorLoss -- with every per-sample host loop of the reference replaced by the batched device
  entry points.  The image regressor (HMR) and the SPIN model are the caller's nn.Modules: they are
  not part of this path and are only called.
"""
import os.path as osp
import pickle

import torch

from .. import ops
from ..utils.geometry import estimate_translation, perspective_projection, rotation_matrix_to_angle_axis
from .fits_dict import FitsDict

_TOPO = {}


def contact_from_verts(verts, contactlists, mode='regions'):
    """[B, n_pairs] minimum squared (expansion-form) distance between the two vertex regions of every
    class in contactlists = {'classes': [(regionA, regionB)], 'csig': {region: vertex ids}}."""
    if mode != 'regions':
        return None
    key = (id(contactlists), verts.shape[1], verts.device)
    topo = _TOPO.get(key)
    if topo is None:
        topo = ops.Topology(torch.zeros(0, 3, dtype=torch.long), verts.shape[1], verts.device)
        topo.set_regions(contactlists)
        _TOPO.clear()                                  # one live region table is all a training run needs
        _TOPO[key] = (topo, contactlists)              # keeps the keyed dict alive (its id cannot be recycled)
        topo = _TOPO[key]
    return topo[0].region_min(verts.detach(), masked=False)[0]


class ContactFromVertsMixin:
    """Drop-in method for a TUCH-like module that owns `self.contactlists` (train_module.py:65-67)."""

    def contact_from_verts(self, verts, mode='regions'):
        return contact_from_verts(verts, self.contactlists, mode)


class TUCH(ContactFromVertsMixin):
    """Same constructor arguments as the reference (train_module.py:31-67).  The three things the
    reference reads from its un-shipped data tree can be handed in instead: fits_dict= (else
    FitsDict(options, train_ds)), contactlists= (else classes.pkl / ContactSigSMPL.pkl under
    config.DSC_ROOT) and focal_length= (else constants.FOCAL_LENGTH)."""

    def __init__(self, options, device, datasets, bodymodel, spin_model, regressor, optimization, criterion,
                 geodistssmpl, fits_dict=None, contactlists=None, focal_length=None, geothres=None, euclthres=None):
        self.options = options
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise ops.TuchError('TUCH needs a CUDA device: tuch_b200 has no CPU fallback')
        if focal_length is None:
            try:
                from data.essentials import constants
                focal_length = constants.FOCAL_LENGTH
            except Exception as e:
                raise ops.TuchError('TUCH: focal_length= not given and data.essentials.constants is not '
                                    'importable: %s' % (e,))
        self.focal_length = focal_length
        self.train_ds, self.val_ds = datasets if datasets is not None else (None, None)
        self.fits_dict = fits_dict if fits_dict is not None else FitsDict(self.options, self.train_ds, device=device)
        self.modelspin = spin_model
        self.model = regressor
        self.smplify = optimization
        self.smpl = bodymodel
        self.geodistssmpl = geodistssmpl
        self.criterion_cospin = criterion
        if contactlists is None:
            try:
                from configs import config
                with open(osp.join(config.DSC_ROOT, 'classes.pkl'), 'rb') as f:
                    classes = pickle.load(f)
                with open(osp.join(config.DSC_ROOT, 'ContactSigSMPL.pkl'), 'rb') as f:
                    csig = pickle.load(f)
            except Exception as e:
                raise ops.TuchError('TUCH: contactlists= not given and the DSC region files are not readable: %s' % (e,))
            contactlists = {'classes': classes, 'csig': csig}
        self.contactlists = contactlists
        self._geothres, self._euclthres = geothres, euclthres
        self._contact_topo = None

    # ------------------------------------------------------------------ train_module.py:93-110
    def get_verts_in_contact(self, verts):
        """{body: [rows with a partner closer than euclthres and at least geothres away along the
        surface, that partner]} (config.geothres / config.euclthres unless given to the constructor)."""
        geothres, euclthres = self._geothres, self._euclthres
        if geothres is None or euclthres is None:
            from configs import config
            geothres = config.geothres if geothres is None else geothres
            euclthres = config.euclthres if euclthres is None else euclthres
        if self._contact_topo is None:
            topo = ops.Topology(torch.zeros(0, 3, dtype=torch.long), verts.shape[1], verts.device)
            topo.set_geomask(self.geodistssmpl >= geothres)
            self._contact_topo = topo
        q = self._contact_topo.contact_query(verts.detach(), use_segments=False, want_winding=False)
        close = q['min_sq'] < euclthres ** 2
        return {b: [torch.where(close[b])[0], q['argmin'][b][close[b]].long()] for b in range(verts.shape[0])}

    # ------------------------------------------------------------------ train_module.py:112-336
    def forward_train_step(self, input_batch):
        o = self.options
        camera_center = torch.zeros(o.batch_size, 2, device=self.device)
        self.model.train()

        images = input_batch['img']
        batch_size = images.shape[0]
        indices = input_batch['sample_index']
        is_flipped = input_batch['is_flipped']
        rot_angle = input_batch['rot_angle']
        dataset_name = input_batch['dataset_name']

        has_pose_3d = input_batch['has_pose_3d'].bool()
        has_disc_contact = input_batch['has_disc_contact'].bool()
        has_2d_keypoints_gtanno = input_batch['has_gt_kpts'].bool()
        has_smpl_ = input_batch['has_smpl'].bool() | input_batch['has_pgt_smpl'].bool()

        gt_keypoints_2d = input_batch['keypoints']
        gt_joints = input_batch['pose_3d']
        gt_pose = input_batch['pose']
        gt_betas = input_batch['betas']
        gt_disc_contact = input_batch['contact_vec']
        gt_out = self.smpl(betas=gt_betas, body_pose=gt_pose[:, 3:], global_orient=gt_pose[:, :3])
        gt_model_joints, gt_verts = gt_out.joints, gt_out.vertices

        # keypoints from [-1,1] to pixels                                                    :148-151
        gt_keypoints_2d_orig = gt_keypoints_2d.clone()
        gt_keypoints_2d_orig[:, :, :-1] = 0.5 * o.img_res * (gt_keypoints_2d_orig[:, :, :-1] + 1)

        # current best fits                                                                  :156-166
        opt_pose, opt_betas = self.fits_dict[(dataset_name, indices.cpu(), rot_angle.cpu(), is_flipped.cpu())]
        opt_pose, opt_betas = opt_pose.to(self.device), opt_betas.to(self.device)
        opt_output = self.smpl(betas=opt_betas, body_pose=opt_pose[:, 3:], global_orient=opt_pose[:, :3])
        opt_vertices, opt_joints = opt_output.vertices, opt_output.joints
        opt_contact_l3 = self.contact_from_verts(opt_vertices, mode='regions')

        # camera translations (one kernel per call, no host round trip)                      :171-183
        gt_cam_t = estimate_translation(gt_model_joints, gt_keypoints_2d_orig, focal_length=self.focal_length,
                                        img_size=o.img_res, has_2d_kp_anno=has_2d_keypoints_gtanno)
        opt_cam_t = estimate_translation(opt_joints, gt_keypoints_2d_orig, focal_length=self.focal_length,
                                         img_size=o.img_res, has_2d_kp_anno=has_2d_keypoints_gtanno)
        center = 0.5 * o.img_res * torch.ones(batch_size, 2, device=self.device)
        opt_joint_loss = self.smplify.get_fitting_loss(opt_pose, opt_betas, opt_cam_t, center, gt_keypoints_2d_orig,
                                                       has_2d_keypoints_gtanno).mean(dim=-1)

        # SPIN fits, for logging only                                                        :186-196
        spin_vertices = spin_cam_t = None
        if self.modelspin is not None:
            with torch.no_grad():
                rot_s, betas_s, cam_s = self.modelspin(images)
                spin_vertices = self.smpl(betas=betas_s, body_pose=rot_s[:, 1:], global_orient=rot_s[:, 0].unsqueeze(1),
                                          pose2rot=False).vertices.clone()
                spin_cam_t = torch.stack([cam_s[:, 1], cam_s[:, 2],
                                          2 * self.focal_length / (o.img_res * cam_s[:, 0] + 1e-9)], dim=-1)

        # regressor                                                                          :202-229
        pred_rotmat, pred_betas, pred_camera = self.model(images)
        pred_output = self.smpl(betas=pred_betas, body_pose=pred_rotmat[:, 1:],
                                global_orient=pred_rotmat[:, 0].unsqueeze(1), pose2rot=False)
        pred_vertices, pred_joints = pred_output.vertices, pred_output.joints
        pred_pose = rotation_matrix_to_angle_axis(pred_rotmat.detach().reshape(-1, 3, 3)).contiguous() \
            .view(batch_size, -1)
        pred_pose[torch.isnan(pred_pose)] = 0.0
        pred_cam_t = torch.stack([pred_camera[:, 1], pred_camera[:, 2],
                                  2 * self.focal_length / (o.img_res * pred_camera[:, 0] + 1e-9)], dim=-1)
        pred_keypoints_2d = perspective_projection(
            pred_joints, rotation=torch.eye(3, device=self.device).unsqueeze(0).expand(batch_size, -1, -1),
            translation=pred_cam_t, focal_length=self.focal_length, camera_center=camera_center)
        pred_keypoints_2d = pred_keypoints_2d / (o.img_res / 2.)

        # SMPLify-DC in the loop, starting from the prediction                              :236-284
        smplifyoptiverts = None
        if o.run_smplify:
            new_opt_vertices, new_opt_joints, new_opt_pose, new_opt_betas, new_opt_cam_t, new_opt_joint_loss, \
                smplifyoptiverts = self.smplify(
                    pred_pose.detach(), pred_betas.detach(), pred_cam_t.detach(), center, gt_keypoints_2d_orig,
                    use_contact=o.use_contact_in_the_loop, contactlist=self.contactlists,
                    gt_contact=[gt_disc_contact, None], ignore_idxs=has_smpl_,
                    has_discrete_contact=has_disc_contact, has_gt_keypoints=has_2d_keypoints_gtanno,
                    contact_loss_weight=o.contact_in_the_loop_loss_weight, contact_loss_return='sum',
                    segments=self.criterion_cospin.segments)
            new_opt_joint_loss = new_opt_joint_loss.mean(dim=-1)
            update = (new_opt_joint_loss <= opt_joint_loss)
            # with discrete contact labels the new fit must also bring the labelled regions at least as close
            new_opt_contact_l3 = self.contact_from_verts(new_opt_vertices, mode='regions')
            update_contact_l3 = ((gt_disc_contact * new_opt_contact_l3) <= (gt_disc_contact * opt_contact_l3)).sum(1) > 0
            if o.use_contact_in_the_loop:
                update[has_disc_contact] = (update_contact_l3 * update)[has_disc_contact]

            opt_joint_loss[update] = new_opt_joint_loss[update]
            opt_vertices[update, :] = new_opt_vertices[update, :]
            opt_contact_l3[update, :] = new_opt_contact_l3[update, :]
            opt_joints[update, :] = new_opt_joints[update, :]
            opt_pose[update, :] = new_opt_pose[update, :]
            opt_betas[update, :] = new_opt_betas[update, :]
            opt_cam_t[update, :] = new_opt_cam_t[update, :]
            self.fits_dict[(dataset_name, indices.cpu(), rot_angle.cpu(), is_flipped.cpu(), update.cpu())] = \
                (opt_pose.cpu(), opt_betas.cpu())

        # ground-truth parameters win where they exist                                       :290-301
        opt_cam_t[has_smpl_, :] = gt_cam_t[has_smpl_, :]
        opt_joints[has_smpl_, :, :] = gt_model_joints[has_smpl_, :, :]
        opt_pose[has_smpl_, :] = gt_pose[has_smpl_, :]
        opt_betas[has_smpl_, :] = gt_betas[has_smpl_, :]
        opt_vertices[has_smpl_, :] = gt_verts[has_smpl_, :]
        valid_fit = (opt_joint_loss < o.smplify_threshold).to(self.device)
        valid_fit_pose = has_smpl_ | valid_fit
        valid_fit_shape = has_smpl_ | valid_fit

        loss, loss_dict = self.criterion_cospin(pred_rotmat, pred_betas, opt_pose, opt_betas, pred_keypoints_2d,
                                                gt_keypoints_2d, pred_joints, gt_joints, has_pose_3d, pred_vertices,
                                                opt_vertices, pred_camera, valid_fit_pose, valid_fit_shape)
        losses = {'loss': loss.detach()}
        for k, val in loss_dict.items():
            losses[k] = val.detach()
        output = {'pred_vertices': pred_vertices.detach(), 'spin_vertices': spin_vertices,
                  'opt_vertices': opt_vertices.detach(), 'pred_cam_t': pred_cam_t.detach(), 'spin_cam_t': spin_cam_t,
                  'opt_cam_t': opt_cam_t.detach(), 'smplifyoptiverts': smplifyoptiverts,
                  'gt_contact_l3': gt_disc_contact, 'has_contact_pc': has_disc_contact,
                  'has_contact': has_disc_contact, 'valid_kpts_anno': valid_fit | has_smpl_,
                  'gt_keypoints': gt_keypoints_2d_orig}
        return loss, losses, output
