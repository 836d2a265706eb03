"""Host-side mirror of tuch/utils/geometry.py: batch_rodrigues (:29-43), quat_to_rotmat (:45-65),
rot6d_to_rotmat (:67-81), perspective_projection (:83-111), estimate_translation (:156-205), plus
torchgeometry's rotation_matrix_to_angle_axis as called at tuch/train/train_module.py:208-211.

These are the small glue functions around the hot path (a few dozen floats per body); they are
expressed with torch tensor ops on the caller's device and are differentiable.  Inside the
SMPLify-DC objective the projection never runs through here: it is fused with the GMoF
reprojection term in one kernel (tuch_reprojection_loss, see tuch_b200/smplify/losses.py).
"""
import torch


def quat_to_rotmat(quat):
    """[B,4] (w, x, y, z), not necessarily unit -> [B,3,3]."""
    q = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q.unbind(dim=1)
    ww, xx, yy, zz = w * w, x * x, y * y, z * z
    rows = [ww + xx - yy - zz, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z,
            2 * w * z + 2 * x * y, ww - xx + yy - zz, 2 * y * z - 2 * w * x,
            2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, ww - xx - yy + zz]
    return torch.stack(rows, dim=1).view(-1, 3, 3)


def batch_rodrigues(theta):
    """Axis-angle [B,3] -> rotation matrices [B,3,3] through the unit quaternion
    (cos(a/2), sin(a/2) * axis), a = |theta + 1e-8|."""
    angle = torch.norm(theta + 1e-8, p=2, dim=1, keepdim=True)
    axis = theta / angle
    half = 0.5 * angle
    return quat_to_rotmat(torch.cat([torch.cos(half), torch.sin(half) * axis], dim=1))


def rot6d_to_rotmat(x):
    """Zhou et al. 6-D rotation representation [B,6] (or [B*k,6]) -> [B,3,3] by Gram-Schmidt on
    the two columns of x.view(-1,3,2)."""
    x = x.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = torch.nn.functional.normalize(a1, dim=1)
    b2 = torch.nn.functional.normalize(a2 - (b1 * a2).sum(dim=1, keepdim=True) * b1, dim=1)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack((b1, b2, b3), dim=-1)


def perspective_projection(points, rotation, translation, focal_length, camera_center):
    """points[bs,N,3], rotation[bs,3,3], translation[bs,3], focal_length scalar or [bs],
    camera_center[bs,2] -> pixel coordinates [bs,N,2] of K (R p + t) / z."""
    p = torch.matmul(points, rotation.transpose(1, 2)) + translation.unsqueeze(1)
    p = p / p[:, :, 2:3]
    f = focal_length
    if isinstance(f, torch.Tensor) and f.dim() > 0:
        f = f.view(-1, 1, 1)
    return f * p[:, :, :2] + camera_center.unsqueeze(1)


def estimate_translation(S, joints_2d, focal_length=5000., img_size=224., has_2d_kp_anno=None):
    """Camera translation that brings the 3-D joints S[B,49,3] closest to joints_2d[B,49,3] (x, y, conf):
    GT joints 25:49 where has_2d_kp_anno, OpenPose joints 0:25 elsewhere -> [B,3] on S's device.
    The reference loops over the batch with a device->host copy and a numpy solve per body
    (geometry.py:191-203); here the whole batch is one kernel (tuch_estimate_translation)."""
    from .. import ops
    if has_2d_kp_anno is None:
        raise ops.TuchError('estimate_translation: has_2d_kp_anno is required (geometry.py:192 indexes it)')
    return ops.estimate_translation(S, joints_2d, has_2d_kp_anno, focal_length, img_size)


def rotation_matrix_to_angle_axis(rotation_matrix):
    """torchgeometry.rotation_matrix_to_angle_axis: [N,3,4] (or [N,3,3]) -> [N,3] (tuch_rotmat_to_angle_axis)."""
    from .. import ops
    return ops.rotmat_to_angle_axis(rotation_matrix)
