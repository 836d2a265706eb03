"""TEST INFRASTRUCTURE -- torch-CPU restatement of the SMPLify-DC objective:
tuch/smplify/losses.py, tuch/smplify/prior.py (merged max-mixture), tuch/utils/geometry.py:83-111.

The V x V and Q x F pair loops run in the C oracle (oracle.clib); everything differentiable is
plain torch so autograd supplies reference gradients.  dtype follows the inputs (fp32 / fp64).
"""
import math
import numpy as np
import torch

from . import clib


def _np_dtype(t):
    return np.float64 if t.dtype == torch.float64 else np.float32


# ---------------------------------------------------------------- geometry.py:83-111
def project(points, translation, focal_length, camera_center):
    """perspective_projection with identity rotation (the only way the hot path calls it,
    losses.py:56-59): K ((p + t) / z), xy only."""
    p = points + translation[:, None, :]
    p = p / p[:, :, 2:3]
    return focal_length * p[:, :, :2] + camera_center[:, None, :]


# ---------------------------------------------------------------- losses.py:25-32
def gmof(x, sigma):
    x2, s2 = x * x, sigma * sigma
    return s2 * x2 / (s2 + x2)


def reprojection_term(joints, cam_t, center, joints_2d, conf, focal_length, sigma=100.0):
    """[B,49]: conf^2 * sum_xy gmof(proj - target)   (losses.py:56-61, 138-143, 178-183)."""
    err = gmof(project(joints, cam_t, focal_length, center) - joints_2d, sigma)
    return conf * conf * err.sum(-1)


# ---------------------------------------------------------------- prior.py:36-132
class GMMPrior:
    """MaxMixturePrior(num_gaussians=8, use_merged=True): buffers as prior.py:80-96."""

    def __init__(self, gmm, dtype=torch.float32):
        covs = np.asarray(gmm['covars'], np.float64)
        self.means = torch.tensor(np.asarray(gmm['means']), dtype=dtype)
        npd = np.float64 if dtype == torch.float64 else np.float32
        prec = np.stack([np.linalg.inv(c) for c in covs.astype(npd)]).astype(npd)     # prior.py:82-83
        self.precisions = torch.tensor(prec, dtype=dtype)
        sqrdets = np.array([np.sqrt(np.linalg.det(c)) for c in covs])                # prior.py:89-90
        const = (2 * np.pi) ** (69 / 2.0)
        nllw = np.asarray(gmm['weights']) / (const * (sqrdets / sqrdets.min()))      # prior.py:93-94
        self.nll_weights = torch.tensor(nllw, dtype=dtype)[None]

    def __call__(self, pose, betas=None):
        diff = pose[:, None, :] - self.means                                         # [B,M,69]
        quad = (torch.einsum('mij,bmj->bmi', self.precisions, diff) * diff).sum(-1)
        ll = 0.5 * quad - torch.log(self.nll_weights)                                # prior.py:124-125
        return ll.min(dim=1)[0]


# ---------------------------------------------------------------- losses.py:125-152
def camera_fitting_loss(joints, betas, cam_t, cam_t_est, center, joints_2d, conf,
                        focal_length=5000.0, depth_loss_weight=100.0, sigma=100.0,
                        shape_prior_weight=0.0):
    rep = reprojection_term(joints, cam_t, center, joints_2d, conf, focal_length, sigma)
    depth = depth_loss_weight ** 2 * (cam_t[:, 2] - cam_t_est[:, 2]) ** 2
    shape = shape_prior_weight ** 2 * (betas ** 2).sum(-1)
    return (rep.sum(-1) + depth + shape).sum()


# ---------------------------------------------------------------- losses.py:155-197
def angle_prior(pose):
    sel = pose[:, [52, 55, 9, 12]] * torch.tensor([1.0, -1.0, -1.0, -1.0], dtype=pose.dtype)
    return torch.exp(sel) ** 2


def body_fitting_loss(body_pose, betas, joints, cam_t, center, joints_2d, conf, prior,
                      focal_length=5000.0, sigma=100.0, pose_prior_weight=4.78,
                      shape_prior_weight=5.0, angle_prior_weight=15.2, output='sum'):
    rep = reprojection_term(joints, cam_t, center, joints_2d, conf, focal_length, sigma)
    if output == 'reprojection':
        return rep
    total = (rep.sum(-1) + pose_prior_weight ** 2 * prior(body_pose, betas)
             + angle_prior_weight ** 2 * angle_prior(body_pose).sum(-1)
             + shape_prior_weight ** 2 * (betas ** 2).sum(-1))
    return total.sum()


# ---------------------------------------------------------------- losses.py:73-117 (per body)
def contact_query(verts_b, faces, geomask_np, segments=None, always_segments=False):
    """No-grad part for ONE body: exterior[V] bool (winding <= 0.99, with the allowed
    self-intersection whitelist), argmin[V], masked min of P, winding[V].
    losses.py:79-93 / loss.py:251-270."""
    dt = _np_dtype(verts_b)
    v = verts_b.detach().numpy().astype(dt)
    tris = v[faces]
    wn = clib.winding_numbers(v, tris, dtype=dt)
    exterior = wn <= 0.99
    if segments is not None and (always_segments or (~exterior).sum() > 0):
        for seg in segments:
            seg_ext = seg.exterior(v, dtype=dt)                     # segmentation.py:81-99
            exterior[seg.vidx[~seg_ext]] = True                    # losses.py:88-89
    am, mn = clib.masked_nearest(v, geomask_np, dtype=dt)
    return exterior, am, mn, wn


def push_pull(verts_b, argmin, exterior, euclthres):
    """losses.py:96-105 for one body: sum tanh^2 push over interior vertices + pull over
    exterior vertices closer than euclthres to their geodesically-far nearest vertex."""
    am = torch.as_tensor(argmin, dtype=torch.long)
    ext = torch.as_tensor(exterior)
    d = torch.norm(verts_b - verts_b[am], dim=1)
    push = (torch.tanh(d[~ext] / 0.04) ** 2).sum() if (~ext).any() else d.new_zeros(())
    near = ext & (d < euclthres)
    pull = (0.005 * torch.tanh(d[near] / 0.005) ** 2).sum() if near.any() else d.new_zeros(())
    return push + pull


def r2r_term(verts_b, geomask_np, cdict, active):
    """losses.py:108-117 for one body: sum over annotated pairs of the min masked squared
    distance between the two regions; differentiable through the attaining pair, in the
    expansion form of contact.py:42 (|x|^2 + |y|^2 - 2 x.y)."""
    dt = _np_dtype(verts_b)
    v = verts_b.detach().numpy().astype(dt)
    total = verts_b.new_zeros(())
    for ci in active:
        ra, rb = cdict['classes'][int(ci)]
        ia, ib = np.asarray(cdict['csig'][ra]), np.asarray(cdict['csig'][rb])
        mn, pa, pb = clib.region_min(v, geomask_np, ia, ib, dtype=dt)
        i, j = int(ia[pa]), int(ib[pb])
        if math.isinf(mn):
            total = total + float('inf')
            continue
        x, y = verts_b[i], verts_b[j]
        total = total + ((x * x).sum() + (y * y).sum() - 2 * (x * y).sum())
    return total


def contact_fitting_loss(body_pose, betas, joints, geomask_np, euclthres, cam_t, center,
                         joints_2d, conf, prior, cdict, gt_contact, ignore_idxs,
                         has_discrete_contact, verts, faces, focal_length=5000.0, sigma=100.0,
                         pose_prior_weight=1.0, contact_loss_weight=1000.0, segments=None,
                         return_parts=False):
    """losses.py:34-123.  faces: numpy [F,3]; geomask_np: numpy bool [V,V];
    segments: list of oracle.segments.Segment or None."""
    B = body_pose.shape[0]
    rep = reprojection_term(joints, cam_t, center, joints_2d, conf, focal_length, sigma)
    pri = pose_prior_weight ** 2 * prior(body_pose, betas)
    contact, r2r = [], []
    aux = []
    for b in range(B):
        if bool(ignore_idxs[b]):
            contact.append(verts.new_zeros(()))
            r2r.append(verts.new_zeros(()))
            aux.append(None)
            continue
        ext, am, mn, wn = contact_query(verts[b], faces, geomask_np, segments)
        contact.append(push_pull(verts[b], am, ext, euclthres))
        if bool(has_discrete_contact[b]):
            active = np.where(np.asarray(gt_contact[b]) == 1)[0]
            r2r.append(r2r_term(verts[b], geomask_np, cdict, active))
        else:
            r2r.append(verts.new_zeros(()))
        aux.append((ext, am, mn, wn))
    contact, r2r = torch.stack(contact), torch.stack(r2r)
    total = rep.sum(-1) + 10 * contact + pri + contact_loss_weight * r2r          # losses.py:120
    if return_parts:
        return total.sum(), dict(reprojection=rep, prior=pri, contact=contact, r2r=r2r, aux=aux)
    return total.sum()
