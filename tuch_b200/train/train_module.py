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
    # The step is written as five stages over a small per-step record; the statements they replace are cited.
    class _Fit:
        """The best-known body of every sample in the batch (the `opt_*` values of the reference)."""
        __slots__ = ('pose', 'betas', 'vertices', 'joints', 'cam_t', 'joint_loss', 'contact')

    def _body(self, pose, betas):
        out = self.smpl(betas=betas, body_pose=pose[:, 3:], global_orient=pose[:, :3])
        return out.vertices, out.joints

    def _camera_from_weak_perspective(self, cam):
        """(s, tx, ty) -> (tx, ty, 2 f / (res s))                                      :193-196, :213-216"""
        depth = 2 * self.focal_length / (self.options.img_res * cam[:, 0] + 1e-9)
        return torch.stack([cam[:, 1], cam[:, 2], depth], dim=-1)

    def _stored_fits(self, key, kp_px, has_gt_kp, center):
        """Fits-store lookup, their bodies, region distances, camera and reprojection score   :156-183"""
        fit = self._Fit()
        pose, betas = self.fits_dict[key]
        fit.pose, fit.betas = pose.to(self.device), betas.to(self.device)
        fit.vertices, fit.joints = self._body(fit.pose, fit.betas)
        fit.contact = self.contact_from_verts(fit.vertices, mode='regions')
        fit.cam_t = estimate_translation(fit.joints, kp_px, focal_length=self.focal_length,
                                         img_size=self.options.img_res, has_2d_kp_anno=has_gt_kp)
        fit.joint_loss = self.smplify.get_fitting_loss(fit.pose, fit.betas, fit.cam_t, center, kp_px,
                                                       has_gt_kp).mean(dim=-1)
        return fit

    def _fit_in_the_loop(self, fit, key, start, kp_px, center, labels, flags):
        """SMPLify-DC from the network's prediction; samples whose new fit scores at least as well -- and, with
        contact labels, brings the labelled regions at least as close -- replace the stored one      :236-284"""
        o = self.options
        has_dc, has_gt_kp, has_smpl = flags
        verts, joints, pose, betas, cam_t, joint_loss, optiverts = self.smplify(
            *start, center, kp_px, use_contact=o.use_contact_in_the_loop, contactlist=self.contactlists,
            gt_contact=[labels, None], ignore_idxs=has_smpl, has_discrete_contact=has_dc,
            has_gt_keypoints=has_gt_kp, contact_loss_weight=o.contact_in_the_loop_loss_weight,
            contact_loss_return='sum', segments=self.criterion_cospin.segments)
        joint_loss = joint_loss.mean(dim=-1)
        better = joint_loss <= fit.joint_loss
        contact = self.contact_from_verts(verts, mode='regions')
        closer = ((labels * contact) <= (labels * fit.contact)).sum(1) > 0
        if o.use_contact_in_the_loop:
            better[has_dc] = (closer * better)[has_dc]
        for name, new in (('joint_loss', joint_loss), ('vertices', verts), ('contact', contact), ('joints', joints),
                          ('pose', pose), ('betas', betas), ('cam_t', cam_t)):
            getattr(fit, name)[better] = new[better]
        self.fits_dict[key + (better.cpu(),)] = (fit.pose.cpu(), fit.betas.cpu())
        return optiverts

    def forward_train_step(self, input_batch):
        o = self.options
        self.model.train()
        images = input_batch['img']
        n = images.shape[0]
        key = (input_batch['dataset_name'], input_batch['sample_index'].cpu(), input_batch['rot_angle'].cpu(),
               input_batch['is_flipped'].cpu())
        has_pose_3d, has_dc, has_gt_kp = (input_batch[k].bool() for k in ('has_pose_3d', 'has_disc_contact',
                                                                          'has_gt_kpts'))
        has_smpl = input_batch['has_smpl'].bool() | input_batch['has_pgt_smpl'].bool()
        labels = input_batch['contact_vec']
        gt_pose, gt_betas = input_batch['pose'], input_batch['betas']

        # ---- ground truth (or mimicked) bodies, keypoints in pixels                                    :142-151
        gt_vertices, gt_joints_model = self._body(gt_pose, gt_betas)
        kp_px = input_batch['keypoints'].clone()
        kp_px[:, :, :-1] = 0.5 * o.img_res * (kp_px[:, :, :-1] + 1)
        center = 0.5 * o.img_res * torch.ones(n, 2, device=self.device)

        # ---- the stored fits; the camera of the ground-truth body is estimated before get_fitting_loss
        #      zeroes the ignored confidences in kp_px (the order of :171-183)
        gt_cam_t = estimate_translation(gt_joints_model, kp_px, focal_length=self.focal_length, img_size=o.img_res,
                                        has_2d_kp_anno=has_gt_kp)
        fit = self._stored_fits(key, kp_px, has_gt_kp, center)

        # ---- SPIN's own prediction, for logging only                                                   :186-196
        spin_vertices = spin_cam_t = None
        if self.modelspin is not None:
            with torch.no_grad():
                rot_s, betas_s, cam_s = self.modelspin(images)
                spin_vertices = self.smpl(betas=betas_s, body_pose=rot_s[:, 1:], global_orient=rot_s[:, 0].unsqueeze(1),
                                          pose2rot=False).vertices.clone()
                spin_cam_t = self._camera_from_weak_perspective(cam_s)

        # ---- the regressor                                                                             :202-229
        pred_rotmat, pred_betas, pred_camera = self.model(images)
        pred = self.smpl(betas=pred_betas, body_pose=pred_rotmat[:, 1:], global_orient=pred_rotmat[:, 0].unsqueeze(1),
                         pose2rot=False)
        pred_pose = rotation_matrix_to_angle_axis(pred_rotmat.detach().reshape(-1, 3, 3)).contiguous().view(n, -1)
        pred_pose[torch.isnan(pred_pose)] = 0.0
        pred_cam_t = self._camera_from_weak_perspective(pred_camera)
        pred_kp = perspective_projection(pred.joints, rotation=torch.eye(3, device=self.device).expand(n, -1, -1),
                                         translation=pred_cam_t, focal_length=self.focal_length,
                                         camera_center=torch.zeros(o.batch_size, 2, device=self.device))
        pred_kp = pred_kp / (o.img_res / 2.)

        # ---- optimisation in the loop                                                                  :236-284
        optiverts = None
        if o.run_smplify:
            optiverts = self._fit_in_the_loop(fit, key, (pred_pose.detach(), pred_betas.detach(), pred_cam_t.detach()),
                                              kp_px, center, labels, (has_dc, has_gt_kp, has_smpl))

        # ---- ground truth wins where it exists; which fits may supervise the regressor                 :290-301
        for name, gt in (('cam_t', gt_cam_t), ('joints', gt_joints_model), ('pose', gt_pose), ('betas', gt_betas),
                         ('vertices', gt_vertices)):
            getattr(fit, name)[has_smpl] = gt[has_smpl]
        good_fit = (fit.joint_loss < o.smplify_threshold).to(self.device)
        supervise = has_smpl | good_fit

        # ---- losses and the logging dictionaries                                                       :303-336
        loss, parts = self.criterion_cospin(pred_rotmat, pred_betas, fit.pose, fit.betas, pred_kp,
                                            input_batch['keypoints'], pred.joints, input_batch['pose_3d'], has_pose_3d,
                                            pred.vertices, fit.vertices, pred_camera, supervise, supervise)
        losses = {'loss': loss.detach(), **{k: v.detach() for k, v in parts.items()}}
        output = dict(pred_vertices=pred.vertices.detach(), spin_vertices=spin_vertices,
                      opt_vertices=fit.vertices.detach(), pred_cam_t=pred_cam_t.detach(), spin_cam_t=spin_cam_t,
                      opt_cam_t=fit.cam_t.detach(), smplifyoptiverts=optiverts, gt_contact_l3=labels,
                      has_contact_pc=has_dc, has_contact=has_dc, valid_kpts_anno=supervise, gt_keypoints=kp_px)
        return loss, losses, output
