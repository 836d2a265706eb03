"""Host-side mirror of tuch/smplify/losses.py: gmof (:25-32), contact_fitting_loss (:34-123),
camera_fitting_loss (:125-152), angle_prior (:155-162), body_fitting_loss (:164-197) -- same
names, positional order, keyword defaults and return values.

Where the reference loops over the bodies in Python and lets autograd differentiate ~80 ATen
launches per body (three 190 MB distance matrices and a 3.4 GB solid-angle tensor each), every
term here is one sm_100a kernel over the whole batch that writes the value AND its analytic
gradient; the autograd.Function below only hands those gradients back to torch.
"""
import collections

import numpy as np
import torch

from .. import ops
from .prior import MaxMixturePrior


def gmof(x, sigma):
    """Geman-McClure robustifier sigma^2 x^2 / (sigma^2 + x^2) (elementwise, any device)."""
    x2, s2 = x ** 2, sigma ** 2
    return (s2 * x2) / (s2 + x2)


def angle_prior(pose):
    """exp(+-pose[:, [52, 55, 9, 12]])^2: penalises unnatural knee / elbow bending (body pose
    without the global rotation, hence the -3 in the reference's indices)."""
    sign = torch.tensor([1., -1., -1., -1.], device=pose.device, dtype=pose.dtype)
    return torch.exp(pose[:, [55 - 3, 58 - 3, 12 - 3, 15 - 3]] * sign) ** 2


# ------------------------------------------------------------------ topology cache
_TOPO_CACHE = collections.OrderedDict()
_TOPO_CACHE_MAX = 4


def topology_for(geomask, face_tensor, num_verts, cdict=None, segments=None):
    """Device-resident constants for (geomask, faces, cdict, segments), built once per distinct set of
    caller objects (the reference re-derives all of it from these arguments on every call)."""
    if isinstance(geomask, ops.Topology):
        return geomask
    if not isinstance(geomask, torch.Tensor) or not geomask.is_cuda:
        raise ops.TuchError('geomask must be a CUDA bool tensor [V,V]: tuch_b200 has no CPU fallback')
    faces = face_tensor[0] if face_tensor.dim() == 3 else face_tensor
    key = (geomask.data_ptr(), geomask._version, tuple(geomask.shape), faces.data_ptr(), tuple(faces.shape),
           id(cdict) if cdict is not None and len(cdict) else None, id(segments) if segments is not None else None)
    hit = _TOPO_CACHE.get(key)
    if hit is not None:
        _TOPO_CACHE.move_to_end(key)
        return hit[0]
    topo = ops.Topology(faces, num_verts, geomask.device)
    topo.set_geomask(geomask)
    if cdict is not None and len(cdict):
        topo.set_regions(cdict)
    if segments is not None:
        topo.set_segments(segments.topology_entries())
    # the cached entry keeps the keyed objects alive so that ids / pointers cannot be recycled
    _TOPO_CACHE[key] = (topo, geomask, faces, cdict, segments)
    while len(_TOPO_CACHE) > _TOPO_CACHE_MAX:
        _TOPO_CACHE.popitem(last=False)
    return topo


# ------------------------------------------------------------------ fused objective
class _FitObjective(torch.autograd.Function):
    """total = sum_b [ sum_j reprojection[b,j] + depth[b] + pose_terms[b] + 10 contact[b] + w r2r[b] ].

    All values and gradients come out of the kernels in forward(); backward() scales them by the
    upstream scalar."""

    @staticmethod
    def forward(ctx, cfg, joints, cam_t, body_pose, betas, verts):
        dev = joints.device
        B = joints.shape[0]
        need = list(ctx.needs_input_grad[1:])
        rep = ops.reprojection_loss(joints, cam_t, cfg['center'], cfg['joints_2d'], cfg['conf'], cfg['focal'],
                                    cfg['sigma'], cam_t_est=cfg.get('cam_t_est'),
                                    depth_loss_weight=cfg.get('depth_w', 0.0), want_grad=need[0] or need[1])
        per_body = rep['loss'].sum(dim=-1)
        if rep['depth'] is not None:
            per_body = per_body + rep['depth']
        g_pose = g_betas = g_verts = None
        aux = {}
        if cfg.get('prior') is not None or cfg.get('apw', 0.0) or cfg.get('spw', 0.0):
            pt = ops.pose_terms(cfg.get('prior'), body_pose, betas if cfg.get('spw', 0.0) else None,
                                pose_prior_weight=cfg.get('ppw', 0.0) if cfg.get('prior') is not None else 0.0,
                                angle_prior_weight=cfg.get('apw', 0.0), shape_prior_weight=cfg.get('spw', 0.0),
                                want_grad=need[2] or need[3])
            per_body = per_body + pt['value']
            g_pose, g_betas = pt['g_pose'], pt['g_betas']
            aux['prior'] = pt['prior']
        if cfg.get('topo') is not None:
            topo = cfg['topo']
            active = cfg['body_active']
            q = topo.contact_query(verts, use_segments=cfg['use_segments'])
            if need[4]:
                g_verts = torch.zeros(B, topo.V, 3, device=dev, dtype=torch.float32)
            contact, _ = ops.contact_loss(verts, q['argmin'], q['exterior'], cfg['euclthres'],
                                          pull_mode=ops.PULL_THRESHOLD, reduce_mode=ops.REDUCE_SUM,
                                          body_active=active, weight=10.0, g_points=g_verts)
            per_body = per_body + 10.0 * contact
            aux.update(contact=contact, exterior=q['exterior'], argmin=q['argmin'], winding=q['winding'])
            if cfg.get('pair_active') is not None:
                mn, ai, aj = topo.region_min(verts, masked=True, active=cfg['pair_active'])
                r2r = ops.region_sum(verts, mn, ai, aj, body_active=active, weight=cfg['r2r_w'], g_verts=g_verts)
                per_body = per_body + cfg['r2r_w'] * r2r
                aux['r2r'] = r2r
        ctx.save_for_backward(*[t if t is not None else torch.empty(0, device=dev)
                                for t in (rep['g_joints'], rep['g_cam_t'], g_pose, g_betas, g_verts)])
        ctx.have = [t is not None for t in (rep['g_joints'], rep['g_cam_t'], g_pose, g_betas, g_verts)]
        aux['per_body'] = per_body
        cfg['aux'] = aux
        cfg['per_body'] = per_body
        return per_body.sum()

    @staticmethod
    def backward(ctx, g):
        outs = []
        for have, need, t in zip(ctx.have, ctx.needs_input_grad[1:], ctx.saved_tensors):
            outs.append(t * g if (have and need) else None)
        return (None,) + tuple(outs)


def _no_grad_like(t):
    return t.detach() if isinstance(t, torch.Tensor) else t


def _pair_activity(gt_contact_l3, has_discrete_contact, ignore_idxs, n_pairs):
    """[B,n_pairs] bool: annotated (== 1) region pairs of the bodies that take part (losses.py:109-112)."""
    act = (gt_contact_l3 == 1) & has_discrete_contact.bool().view(-1, 1) & (~ignore_idxs.bool()).view(-1, 1)
    if act.shape[1] != n_pairs:
        raise ops.TuchError('gt_contact has %d classes, cdict %d' % (act.shape[1], n_pairs))
    return act


def contact_fitting_loss(body_pose, global_orient, body_pose_loop1, opt_global_orient_smplifyloop1,
                         betas, model_joints, geomask, euclthres,
                         camera_t, camera_center,
                         joints_2d, joints_conf, pose_prior,
                         cdict, gt_contact,
                         ignore_idxs,
                         has_discrete_contact,
                         verts, face_tensor=None,
                         device='cuda',
                         focal_length=5000, sigma=100, pose_prior_weight=1.0,
                         shape_prior_weight=1.0, angle_prior_weight=1.0,
                         contact_loss_weight=1000, output='sum',
                         segments=None, return_parts=False):
    """Loss function for body fitting with contact (losses.py:34-123).  Returns the 0-d total
    sum_b [ sum_j conf^2 gmof(reprojection) + 10 (push + pull) + ppw^2 prior + w r2r ].

    `geomask` is the [V,V] bool tensor of the reference or an ops.Topology that already holds it
    (then face_tensor / cdict / segments are taken from the topology).  global_orient, the *_loop1
    tensors, shape/angle_prior_weight, device and output are accepted and unused, as in the reference."""
    B = body_pose.shape[0]
    if isinstance(geomask, ops.Topology):
        topo = geomask
    else:
        if face_tensor is None:
            raise ops.TuchError('contact_fitting_loss needs face_tensor')
        topo = topology_for(geomask, face_tensor, verts.shape[1], cdict, segments)
    ignore = ignore_idxs if ignore_idxs is not None else torch.zeros(B, dtype=torch.bool, device=verts.device)
    cfg = dict(center=camera_center, joints_2d=joints_2d, conf=joints_conf, focal=float(focal_length),
               sigma=float(sigma), ppw=float(pose_prior_weight), topo=topo, body_active=~ignore.bool(),
               euclthres=float(euclthres), use_segments=segments is not None or (isinstance(geomask, ops.Topology)
                                                                                 and len(topo.segment_names) > 0),
               r2r_w=float(contact_loss_weight))
    gt_l3 = gt_contact[0] if gt_contact is not None else None
    if gt_l3 is not None and has_discrete_contact is not None and len(topo.classes) > 0:
        cfg['pair_active'] = _pair_activity(gt_l3.to(verts.device), has_discrete_contact.to(verts.device),
                                            ignore, len(topo.classes))
    extra = None
    if isinstance(pose_prior, MaxMixturePrior):
        cfg['prior'] = pose_prior._handle(body_pose.device)
    elif pose_prior is not None:
        # a foreign callable: evaluate it with torch autograd, outside the fused objective
        extra = ((pose_prior_weight ** 2) * pose_prior(body_pose, betas)).sum()
    total = _FitObjective.apply(cfg, model_joints, camera_t, body_pose, None, verts)
    if extra is not None:
        total = total + extra
    if return_parts:
        return total, cfg['aux']
    return total


def camera_fitting_loss(smpl_output, camera_t, camera_t_est, camera_center, joints_2d, joints_conf,
                        focal_length=5000, depth_loss_weight=100, sigma=100, shape_prior_weight=0.0):
    """Loss function for camera and betas optimisation (losses.py:125-152):
    sum_b [ sum_j conf^2 gmof(reprojection) + dw^2 (t_z - t_est_z)^2 + spw^2 |betas|^2 ]."""
    joints, betas = smpl_output.joints, smpl_output.betas
    cfg = dict(center=camera_center, joints_2d=joints_2d, conf=joints_conf, focal=float(focal_length),
               sigma=float(sigma), cam_t_est=camera_t_est, depth_w=float(depth_loss_weight),
               spw=float(shape_prior_weight))
    # betas enter twice: through the joints (autograd of the caller's SMPL forward) and the regulariser
    dummy_pose = torch.zeros(joints.shape[0], 69, device=joints.device) if shape_prior_weight else None
    return _FitObjective.apply(cfg, joints, camera_t, dummy_pose, betas if shape_prior_weight else None, None)


def body_fitting_loss(body_pose, betas, model_joints, camera_t, camera_center,
                      joints_2d, joints_conf, pose_prior,
                      focal_length=5000, sigma=100, pose_prior_weight=4.78,
                      shape_prior_weight=5, angle_prior_weight=15.2,
                      output='sum'):
    """SPIN's body fitting loss (losses.py:164-197): reprojection + pose prior + angle prior + shape
    regulariser; output='reprojection' returns the [B,49] reprojection term alone."""
    if output == 'reprojection':
        if model_joints.requires_grad or camera_t.requires_grad:
            return _ReprojectionOnly.apply(model_joints, camera_t, camera_center, joints_2d, joints_conf,
                                           float(focal_length), float(sigma))
        return ops.reprojection_loss(model_joints, camera_t, camera_center, joints_2d, joints_conf,
                                     focal_length, sigma, want_grad=False)['loss']
    if output != 'sum':
        return None
    cfg = dict(center=camera_center, joints_2d=joints_2d, conf=joints_conf, focal=float(focal_length),
               sigma=float(sigma), ppw=float(pose_prior_weight), apw=float(angle_prior_weight),
               spw=float(shape_prior_weight))
    extra = None
    if isinstance(pose_prior, MaxMixturePrior):
        cfg['prior'] = pose_prior._handle(body_pose.device)
    elif pose_prior is not None:
        extra = ((pose_prior_weight ** 2) * pose_prior(body_pose, betas)).sum()
    total = _FitObjective.apply(cfg, model_joints, camera_t, body_pose, betas, None)
    return total if extra is None else total + extra


class _ReprojectionOnly(torch.autograd.Function):
    @staticmethod
    def forward(ctx, joints, cam_t, center, joints_2d, conf, focal, sigma):
        ctx.save_for_backward(joints, cam_t, center, joints_2d, conf)
        ctx.fs = (focal, sigma)
        return ops.reprojection_loss(joints, cam_t, center, joints_2d, conf, focal, sigma, want_grad=False)['loss']

    @staticmethod
    def backward(ctx, g):
        joints, cam_t, center, joints_2d, conf = ctx.saved_tensors
        r = ops.reprojection_loss(joints, cam_t, center, joints_2d, conf, ctx.fs[0], ctx.fs[1], g_loss=g.contiguous())
        return r['g_joints'], r['g_cam_t'], None, None, None, None, None
