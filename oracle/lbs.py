"""TEST INFRASTRUCTURE -- torch-CPU restatement of the SMPL forward the reference gets from the
third-party package smplx==0.1.13 (requirements.txt:13; call sites tuch/models/smpl.py:22-24,38,
46-47), plus the reference's own wrapper tuch/models/smpl.py:37-56.

PARITY UNPINNED for the smplx part: smplx is not vendored in /root/reference and is not
installable here; this file restates its published algorithm (smplx/lbs.py: blend_shapes,
vertices2joints, batch_rodrigues, batch_rigid_transform, lbs; smplx/body_models.py SMPL.forward;
smplx/vertex_joint_selector.py).  Works in fp32 and fp64 and is differentiable (autograd), so
it doubles as the gradient reference for the fused CUDA LBS backward.
"""
import torch


def rodrigues(rot_vecs):
    """smplx.lbs.batch_rodrigues: angle = |r + 1e-8|, R = I + sin K + (1 - cos) K^2.  [N,3] -> [N,3,3]."""
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)           # [N,1]
    axis = rot_vecs / angle
    s, c = torch.sin(angle)[:, :, None], torch.cos(angle)[:, :, None]  # [N,1,1]
    x, y, z = axis[:, 0], axis[:, 1], axis[:, 2]
    o = torch.zeros_like(x)
    K = torch.stack([o, -z, y, z, o, -x, -y, x, o], dim=1).view(-1, 3, 3)
    eye = torch.eye(3, dtype=rot_vecs.dtype).unsqueeze(0)
    return eye + s * K + (1 - c) * torch.bmm(K, K)


def rigid_chain(rot_mats, joints, parents):
    """smplx.lbs.batch_rigid_transform.  rot_mats[B,J,3,3], joints[B,J,3] ->
    posed_joints[B,J,3], rel_transforms[B,J,4,4]."""
    B, J = joints.shape[:2]
    rel = joints.clone()
    rel[:, 1:] = joints[:, 1:] - joints[:, parents[1:]]
    local = torch.zeros(B, J, 4, 4, dtype=joints.dtype)
    local[:, :, :3, :3] = rot_mats
    local[:, :, :3, 3] = rel
    local[:, :, 3, 3] = 1
    chain = [local[:, 0]]
    for k in range(1, J):
        chain.append(torch.matmul(chain[int(parents[k])], local[:, k]))
    G = torch.stack(chain, dim=1)
    posed = G[:, :, :3, 3]
    jh = torch.cat([joints, torch.zeros(B, J, 1, dtype=joints.dtype)], dim=2).unsqueeze(-1)   # [B,J,4,1]
    corr = torch.matmul(G, jh)                                                                 # [B,J,4,1]
    A = G - torch.cat([torch.zeros(B, J, 4, 3, dtype=joints.dtype), corr], dim=3)
    return posed, A


def lbs(model, betas, full_pose, pose2rot=True):
    """smplx.lbs.lbs.  model: dict of torch tensors (v_template[V,3], shapedirs[V,3,L],
    posedirs[207,3V], J_regressor[24,V], lbs_weights[V,24], parents[24]).
    full_pose: [B,72] (pose2rot) or [B,24,3,3].  Returns vertices[B,V,3], posed joints[B,24,3]."""
    B = betas.shape[0]
    dt = betas.dtype
    v_shaped = model['v_template'] + torch.einsum('bl,mkl->bmk', betas, model['shapedirs'])
    J = torch.einsum('bik,ji->bjk', v_shaped, model['J_regressor'])
    eye = torch.eye(3, dtype=dt)
    if pose2rot:
        R = rodrigues(full_pose.reshape(-1, 3)).view(B, -1, 3, 3)
    else:
        R = full_pose.view(B, -1, 3, 3)
    feat = (R[:, 1:] - eye).reshape(B, -1)
    v_posed = v_shaped + torch.matmul(feat, model['posedirs']).view(B, -1, 3)
    posed_joints, A = rigid_chain(R, J, model['parents'])
    T = torch.matmul(model['lbs_weights'].unsqueeze(0).expand(B, -1, -1), A.view(B, -1, 16)).view(B, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=dt)], dim=2)
    verts = torch.matmul(T, vh.unsqueeze(-1))[:, :, :3, 0]
    return verts, posed_joints


def smpl_forward(model, betas, body_pose, global_orient, pose2rot=True):
    """smplx.SMPL.forward (get_skin=True, no transl) followed by tuch/models/smpl.py:44-56:
    45 joints = 24 posed + 21 picked vertices; + 9 regressed (J_regressor_extra) = 54; remap to 49.
    Returns vertices[B,V,3], joints[B,49,3], full_pose."""
    full_pose = torch.cat([global_orient, body_pose], dim=1)
    verts, j24 = lbs(model, betas, full_pose, pose2rot=pose2rot)
    j45 = torch.cat([j24, verts[:, model['extra_vertex_ids']]], dim=1)
    extra = torch.einsum('bik,ji->bjk', verts, model['J_regressor_extra'])      # smpl.py:47
    j54 = torch.cat([j45, extra], dim=1)                                       # smpl.py:48
    return verts, j54[:, model['joint_map']], full_pose                        # smpl.py:49


def to_torch_model(np_model, dtype=torch.float32):
    out = {}
    for k, v in np_model.items():
        t = torch.as_tensor(v)
        out[k] = t.to(dtype) if t.is_floating_point() else t.long()
    return out
