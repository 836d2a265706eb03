"""Host-side mirror of tuch/train/loss.py: batch_face_normals (:30-41), RegressorLoss (:44-317) --
same constructor arguments, forward signature, loss_dict keys and `.segments` attribute
(read by tuch/train/train_module.py:254).

contact_loss (:240-317) is the hot part: the reference loops over the valid bodies in Python, builds
190 MB distance matrices, a 3.4 GB solid-angle tensor and two huge equality matrices per body; here the
whole batch runs through one C-ABI call (tuch_regressor_contact_loss) that keeps every pairwise
quantity on chip and needs no host synchronisation.  The SPIN terms (:170-236) are small
reductions over [B,49,*] / [B,24,3,3] tensors and stay plain torch ops on the caller's device.

Where the reference reads the HD model files (config.HD_MODEL_DIR) and the segment definitions
(data.essentials.segments.smpl.segm_utils) from its un-shipped data tree, the same arrays can be
passed in directly (`hd_regressor=`, `hd_faces=`, `segments=`).
"""
import os.path as osp
import pickle

import numpy as np
import torch
import torch.nn as nn

from .. import distributed as tdist, ops
from ..utils.geometry import batch_rodrigues
from ..utils.segmentation import BatchBodySegment


def masked_mean(rows, mask, dist_mode=None):
    """`rows[mask].mean()` of the reference (loss.py:182,201,211,225-236,317) for per-body rows of equal size.

    dist_mode None: exactly that, on this process's batch (an empty selection gives NaN, as in the reference).
    One process per GPU with the batch sharded over the ranks: the mean runs over a data-dependent subset of the
    GLOBAL batch, so every rank returns its share  local_sum / global_count  (tuch_b200.distributed.
    global_masked_mean: one scalar all-reduce of the counts).  'sum': the shares add up to the single-process value
    and a SUM all-reduce of the gradients gives the single-process gradient.  'mean': the share is multiplied by
    the world size, for gradient reductions that AVERAGE over the ranks (DistributedDataParallel)."""
    if rows.dim() > 1:
        rows = rows.reshape(rows.shape[0], -1).mean(dim=1)
    if dist_mode is None:
        return rows[mask].mean()
    val, _ = tdist.global_masked_mean(rows, mask)
    if dist_mode == 'mean':
        val = val * tdist.world()[1]
    elif dist_mode != 'sum':
        raise ValueError("dist_mode must be None, 'sum' or 'mean'")
    return val


def batch_face_normals(triangles):
    """Unit normals [B,F,3] of triangles [B,F,3,3] (edge01 x edge02, normalised)."""
    n = torch.cross(triangles[:, :, 1] - triangles[:, :, 0], triangles[:, :, 2] - triangles[:, :, 0], dim=2)
    return n / torch.norm(n, 2, dim=2, keepdim=True)


class _RegressorContact(torch.autograd.Function):
    """contact_loss[valid_fit].mean() with the gradient w.r.t. pred_vertices produced by the kernels."""

    @staticmethod
    def forward(ctx, pred_vertices, topo, valid, euclthres, use_hd, dist_mode=None):
        B = pred_vertices.shape[0]
        valid = valid.bool()
        n_valid = valid.sum()                                                  # stays on the device
        scale = 1.0
        if dist_mode is not None:                                              # global count, see masked_mean()
            rank, ws = tdist.world()
            if ws > 1:
                import torch.distributed as dist
                n_valid = n_valid.clone()
                dist.all_reduce(n_valid, op=dist.ReduceOp.SUM)
            scale = float(ws) if dist_mode == 'mean' else 1.0
        g_loss = valid.float() * scale / n_valid.clamp(min=1).float()
        g_verts = torch.zeros(B, topo.V, 3, device=pred_vertices.device, dtype=torch.float32) \
            if ctx.needs_input_grad[0] else None
        per_body = topo.regressor_contact_loss(pred_vertices, valid=valid, euclthres=euclthres, use_hd=use_hd,
                                               g_loss=g_loss, g_verts=g_verts)
        ctx.save_for_backward(g_verts if g_verts is not None else torch.empty(0, device=pred_vertices.device))
        # mean over an empty selection is NaN in the reference (loss.py:317); keep that
        return (per_body * valid.float()).sum() * scale / n_valid.float()

    @staticmethod
    def backward(ctx, g):
        (g_verts,) = ctx.saved_tensors
        return (g_verts * g if g_verts.numel() else None), None, None, None, None, None


class RegressorLoss(nn.Module):
    def __init__(self, options, device, num_verts, faces, geodistssmpl, geothres=0.2, euclthres=0.02,
                 face_tensor=None, use_hd=True, hd_regressor=None, hd_faces=None, segments=None, template=None):
        super().__init__()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise ops.TuchError('RegressorLoss needs a CUDA device: tuch_b200 has no CPU fallback')
        self.options = options
        self.criterion_shape = nn.L1Loss().to(self.device)
        self.criterion_keypoints = nn.MSELoss(reduction='none').to(self.device)
        self.criterion_regr = nn.MSELoss().to(self.device)
        self.faces = faces
        self.nv = num_verts
        self.geodistssmpl = geodistssmpl
        self.geothres = geothres
        self.geomask = geodistssmpl > geothres
        self.euclthres = euclthres
        self.face_tensor = face_tensor
        self.use_hd = use_hd
        self.dist_mode = None            # set_distributed(): count-corrected means over a sharded batch
        if self.use_hd:
            if hd_regressor is None or hd_faces is None:
                try:                                               # the reference's own files (loss.py:81-88)
                    from configs import config
                    hd_regressor = np.load(osp.join(config.HD_MODEL_DIR, 'smpl_neutral_hd_vert_regressor.npy'))
                    with open(osp.join(config.HD_MODEL_DIR, 'smpl_neutral_hd_sample_from_mesh_out.pkl'), 'rb') as f:
                        hd_faces = pickle.load(f)['faces_vert_is_sampled_from']
                except Exception as e:
                    raise ops.TuchError('RegressorLoss(use_hd=True): hd_regressor= / hd_faces= not given and the '
                                        'reference HD model files are not readable: %s' % (e,))
            self.geovec = torch.as_tensor(np.asarray(hd_faces), device=self.device)
            self.geovec_verts = self.face_tensor[0][self.geovec][:, 0]
        if segments is None:
            try:                                                   # loss.py:91
                from data.essentials.segments.smpl import segm_utils as exn
            except Exception as e:
                raise ops.TuchError('RegressorLoss: segments= not given and data.essentials.segments.smpl.segm_utils '
                                    'is not importable: %s' % (e,))
            segments = BatchBodySegment([x for x in exn.segments.keys()], self.face_tensor[0])
        self.segments = segments
        self._topo = ops.Topology(self.face_tensor[0], num_verts, self.device)
        self._topo.set_geomask(self.geomask)
        self._topo.set_segments(self.segments.topology_entries())
        if template is not None:                               # e.g. smpl.v_template: the face / vertex hierarchy of the
            self._topo.set_template(template)                  # hierarchical kernels (else: first body seen)
        if self.use_hd:
            self._topo.set_hd(hd_regressor, hd_faces)

    def set_distributed(self, mode):
        """One process per GPU, the batch sharded over the ranks: every `x[subset].mean()` of this loss (contact
        term :317, SPIN terms :182,201,211,225-236, camera term :138) becomes this rank's share of the mean over
        the GLOBAL batch; see masked_mean().  mode: None (single process), 'sum' or 'mean' (how the caller reduces
        the gradients over the ranks)."""
        if mode not in (None, 'sum', 'mean'):
            raise ValueError("mode must be None, 'sum' or 'mean'")
        self.dist_mode = mode
        return self

    def forward(self, pred_rotmat, pred_betas, opt_pose, opt_betas, pred_keypoints_2d, gt_keypoints_2d,
                pred_joints, gt_joints, has_pose_3d, pred_vertices, opt_vertices, pred_camera, valid_fit,
                valid_fit_shape):
        o = self.options
        loss_contact = torch.tensor(0)
        if o.contact_loss_weight > 0:
            loss_contact = self.contact_loss(pred_vertices, valid_fit)
        contact_loss = o.contact_loss_weight * loss_contact
        loss_regr_pose, loss_regr_betas = self.smpl_losses(pred_rotmat, pred_betas, opt_pose, opt_betas,
                                                           valid_fit, valid_fit_shape)
        loss_keypoints = self.keypoint_loss(pred_keypoints_2d, gt_keypoints_2d, o.openpose_train_weight,
                                            o.gt_train_weight, valid_fit)
        loss_keypoints_3d = self.keypoint_3d_loss(pred_joints, gt_joints, has_pose_3d)
        loss_shape = self.shape_loss(pred_vertices, opt_vertices, valid_fit)
        cam_rows = (torch.exp(-pred_camera[:, 0] * 10)) ** 2
        cam_loss = cam_rows.mean() if self.dist_mode is None else \
            masked_mean(cam_rows, torch.ones_like(cam_rows, dtype=torch.bool), self.dist_mode)
        spin_loss = o.shape_loss_weight * loss_shape + o.keypoint_loss_weight * loss_keypoints + \
            o.keypoint_loss_weight * loss_keypoints_3d + o.pose_loss_weight * loss_regr_pose + \
            o.beta_loss_weight * loss_regr_betas + cam_loss
        total_loss = spin_loss + contact_loss
        loss_dict = {'loss_shape': loss_shape, 'loss_keypoints': loss_keypoints,
                     'loss_keypoints_3d': loss_keypoints_3d, 'loss_regr_pose': loss_regr_pose,
                     'loss_regr_betas': loss_regr_betas, 'loss_cam': cam_loss, 'loss_contact': loss_contact}
        return total_loss, loss_dict

    # ------------------------------------------------------------------ SPIN terms (glue)
    def _zero(self):
        return torch.zeros(1, device=self.device)

    def keypoint_loss(self, pred_keypoints_2d, gt_keypoints_2d, openpose_weight, gt_weight, valid_fit=None):
        """confidence-weighted 2-D keypoint MSE, mean over the valid bodies (:170-182)."""
        conf = gt_keypoints_2d[:, :, -1].unsqueeze(-1).clone()
        conf[:, :25] *= openpose_weight
        conf[:, 25:] *= gt_weight
        loss = (conf * self.criterion_keypoints(pred_keypoints_2d, gt_keypoints_2d[:, :, :-1])).mean(dim=(1, 2))
        return masked_mean(loss, valid_fit, self.dist_mode)

    def keypoint_3d_loss(self, pred_keypoints_3d, gt_keypoints_3d, has_pose_3d):
        """pelvis-centred, confidence-weighted 3-D keypoint MSE over the 24 GT joints (:184-203)."""
        sel = has_pose_3d == 1
        if self.dist_mode is not None:
            pred = pred_keypoints_3d[:, 25:, :]
            conf = gt_keypoints_3d[:, :, -1].unsqueeze(-1)
            gt = gt_keypoints_3d[:, :, :-1]
            gt = gt - ((gt[:, 2, :] + gt[:, 3, :]) / 2)[:, None, :]
            pred = pred - ((pred[:, 2, :] + pred[:, 3, :]) / 2)[:, None, :]
            return self._global_or_zero(conf * self.criterion_keypoints(pred, gt), sel)
        pred = pred_keypoints_3d[:, 25:, :][sel]
        conf = gt_keypoints_3d[:, :, -1].unsqueeze(-1).clone()[sel]
        gt = gt_keypoints_3d[:, :, :-1].clone()[sel]
        if len(gt) == 0:
            return self._zero()
        gt = gt - ((gt[:, 2, :] + gt[:, 3, :]) / 2)[:, None, :]
        pred = pred - ((pred[:, 2, :] + pred[:, 3, :]) / 2)[:, None, :]
        return (conf * self.criterion_keypoints(pred, gt)).mean()

    def _global_or_zero(self, rows, sel):
        """Sharded form of `if len(selection) == 0: return zeros(1)`: the emptiness test is GLOBAL (the count
        all-reduce of masked_mean), so all ranks take the same branch without a host synchronisation."""
        rows = rows.reshape(rows.shape[0], -1).mean(dim=1)
        val, count = tdist.global_masked_mean(rows, sel)
        if self.dist_mode == 'mean':
            val = val * tdist.world()[1]
        return torch.where(count > 0, val, torch.zeros_like(val)).reshape(1)

    def shape_loss(self, pred_vertices, gt_vertices, has_smpl):
        """per-vertex L1 on the bodies with a fit (:205-214)."""
        sel = has_smpl == 1
        if self.dist_mode is not None:
            return self._global_or_zero((pred_vertices - gt_vertices).abs(), sel)
        if int(sel.sum()) == 0:
            return self._zero()
        return self.criterion_shape(pred_vertices[sel], gt_vertices[sel])

    def smpl_losses(self, pred_rotmat, pred_betas, gt_pose, gt_betas, has_smpl_pose, has_smpl_shape):
        """MSE on rotation matrices / betas of the bodies with a fit (:216-238)."""
        sp, ss = has_smpl_pose == 1, has_smpl_shape == 1
        gt_rotmat = batch_rodrigues(gt_pose.view(-1, 3)).view(-1, 24, 3, 3)
        if self.dist_mode is not None:
            return (self._global_or_zero((pred_rotmat - gt_rotmat) ** 2, sp),
                    self._global_or_zero((pred_betas - gt_betas) ** 2, ss))
        loss_pose = self.criterion_regr(pred_rotmat[sp], gt_rotmat[sp]) if int(sp.sum()) > 0 else self._zero()
        loss_betas = self.criterion_regr(pred_betas[ss], gt_betas[ss]) if int(ss.sum()) > 0 else self._zero()
        return loss_pose, loss_betas

    # ------------------------------------------------------------------ the hot term
    def contact_loss(self, pred_vertices, valid_fit):
        """Self-contact push/pull loss on the (HD-resampled) predicted mesh, mean over the valid bodies."""
        return _RegressorContact.apply(pred_vertices, self._topo, valid_fit, float(self.euclthres), bool(self.use_hd),
                                       self.dist_mode)
