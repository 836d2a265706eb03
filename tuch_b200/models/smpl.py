"""Host-side mirror of tuch/models/smpl.py (SMPL :34-56, ModelOutput :30-32) over the fused
sm_100a LBS kernels instead of smplx==0.1.13.

Same constructor (`SMPL(model_path, batch_size=, create_transl=False, gender=)`), attributes
(`.faces`, `.get_num_verts()`, `.J_regressor_extra`, `.joint_map`, nn.Module semantics) and
forward keywords (`betas`, `body_pose`, `global_orient`, `pose2rot`, `return_full_pose`);
differentiable w.r.t. betas / body_pose / global_orient in axis-angle and rotation-matrix mode.
"""
import os
import pickle
import sys
import types
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn

from .. import ops

ModelOutput = namedtuple('ModelOutput',
                         ['vertices', 'joints', 'full_pose', 'betas', 'global_orient', 'body_pose'])

# smplx VertexJointSelector for SMPL: face, feet, left-hand tips, right-hand tips
SMPL_EXTRA_VERTEX_IDS = [332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
                         2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133]


def _to_np(x):
    if hasattr(x, 'todense'):
        x = x.todense()
    if hasattr(x, 'r'):
        x = x.r
    return np.asarray(x)


def _install_chumpy_stub():
    """SMPL pickles reference chumpy.Ch objects; when chumpy is absent a tolerant stand-in that only
    keeps the array payload is enough to read them."""
    try:
        import chumpy  # noqa: F401
        return
    except Exception:
        pass

    class Ch:
        def __setstate__(self, state):
            self.__dict__.update(state if isinstance(state, dict) else {})

        @property
        def r(self):
            return np.asarray(self.__dict__.get('x'))
    for name in ('chumpy', 'chumpy.ch', 'chumpy.reordering', 'chumpy.ch_ops'):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.Ch = Ch
            m.__path__ = []
            sys.modules[name] = m


def load_smpl_arrays(model_path, gender='neutral', num_betas=10):
    """Reads SMPL_{GENDER}.pkl (or .npz) the way smplx.SMPL.__init__ does and returns the arrays the
    kernels need."""
    path = model_path
    if os.path.isdir(path):
        for ext in ('pkl', 'npz'):
            cand = os.path.join(path, 'SMPL_{}.{}'.format(gender.upper(), ext))
            if os.path.exists(cand):
                path = cand
                break
    if not os.path.isfile(path):
        raise ops.TuchError('SMPL model file not found under %r' % (model_path,))
    if path.endswith('.npz'):
        data = dict(np.load(path, allow_pickle=True))
    else:
        _install_chumpy_stub()
        with open(path, 'rb') as f:
            data = pickle.load(f, encoding='latin1')
    v_template = _to_np(data['v_template']).astype(np.float32)
    V = v_template.shape[0]
    shapedirs = _to_np(data['shapedirs'])[:, :, :num_betas].astype(np.float32)
    posedirs = _to_np(data['posedirs']).astype(np.float32)
    if posedirs.shape[0] == V:                                    # [V,3,207] on disk
        posedirs = posedirs.reshape(-1, posedirs.shape[-1]).T
    kin = _to_np(data['kintree_table']).astype(np.int64) if 'kintree_table' in data else None
    parents = kin[0].copy() if kin is not None else _to_np(data['parents']).astype(np.int64)
    parents[0] = -1
    weights = _to_np(data['weights'] if 'weights' in data else data['lbs_weights']).astype(np.float32)
    faces = _to_np(data['f'] if 'f' in data else data['faces']).astype(np.int64)
    return dict(v_template=v_template, shapedirs=shapedirs, posedirs=np.ascontiguousarray(posedirs),
                J_regressor=_to_np(data['J_regressor']).astype(np.float32), lbs_weights=weights,
                parents=parents, faces=faces)


class _SmplFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, betas, pose, is_rotmat):
        handle = module._handle(betas.device)
        verts, joints, ws = handle.forward(betas, pose, is_rotmat)
        ctx.handle, ctx.is_rotmat, ctx.ws = handle, is_rotmat, ws
        ctx.save_for_backward(pose.detach())
        return verts, joints

    @staticmethod
    def backward(ctx, g_verts, g_joints):
        (pose,) = ctx.saved_tensors
        need_b, need_p = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        g_pose, g_betas = ctx.handle.backward(pose, ctx.is_rotmat, ctx.ws, g_verts, g_joints,
                                              need_pose=need_p, need_betas=need_b)
        if g_pose is not None:
            g_pose = g_pose.view_as(pose)
        return None, g_betas, g_pose, None


class SMPL(nn.Module):
    """Extension of the SMPL body model to 49 joints (24 + 21 picked + 9 regressed, remapped)."""

    def __init__(self, model_path=None, batch_size=1, create_transl=False, gender='neutral',
                 num_betas=10, model_arrays=None, J_regressor_extra=None, joint_map=None,
                 extra_vertex_ids=None, **kwargs):
        super().__init__()
        if create_transl:
            raise ops.TuchError('create_transl=True is not part of the TUCH path (every reference call site passes False)')
        arrays = dict(model_arrays) if model_arrays is not None else load_smpl_arrays(model_path, gender, num_betas)
        if J_regressor_extra is None:
            J_regressor_extra = arrays.get('J_regressor_extra')
        if joint_map is None:
            joint_map = arrays.get('joint_map')
        if J_regressor_extra is None or joint_map is None:
            # the reference's own sources (tuch/models/smpl.py:39-40)
            try:
                from configs import config
                from data.essentials import constants
                if J_regressor_extra is None:
                    J_regressor_extra = np.load(config.JOINT_REGRESSOR_TRAIN_EXTRA)
                if joint_map is None:
                    joint_map = [constants.JOINT_MAP[i] for i in constants.JOINT_NAMES]
            except Exception as e:
                raise ops.TuchError('SMPL: J_regressor_extra / joint_map not given and the reference data tree '
                                    '(configs.config, data.essentials.constants) is not importable: %s' % (e,))
        if extra_vertex_ids is None:
            extra_vertex_ids = arrays.get('extra_vertex_ids', SMPL_EXTRA_VERTEX_IDS)
        arrays['J_regressor_extra'] = np.asarray(J_regressor_extra, dtype=np.float32)
        arrays['joint_map'] = np.asarray(joint_map, dtype=np.int64)
        arrays['extra_vertex_ids'] = np.asarray(extra_vertex_ids, dtype=np.int64)
        self._arrays = arrays
        self.batch_size = batch_size
        self.num_betas = int(arrays['shapedirs'].shape[-1])
        self.faces = np.asarray(arrays['faces'])
        self.register_buffer('faces_tensor', torch.tensor(self.faces.astype(np.int64), dtype=torch.long))
        self.register_buffer('v_template', torch.tensor(arrays['v_template'], dtype=torch.float32))
        self.register_buffer('J_regressor_extra', torch.tensor(arrays['J_regressor_extra'], dtype=torch.float32))
        self.joint_map = torch.tensor(arrays['joint_map'], dtype=torch.long)
        # smplx keeps default parameters of size batch_size for omitted arguments
        self.register_buffer('_default_betas', torch.zeros(batch_size, self.num_betas))
        self.register_buffer('_default_body_pose', torch.zeros(batch_size, 69))
        self.register_buffer('_default_global_orient', torch.zeros(batch_size, 3))
        self._handles = {}

    def get_num_verts(self):
        return int(self._arrays['v_template'].shape[0])

    def get_num_faces(self):
        return int(self.faces.shape[0])

    def _handle(self, device):
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        h = self._handles.get(key)
        if h is None:
            h = ops.SmplHandle(self._arrays, device)
            self._handles[key] = h
        return h

    def forward(self, betas=None, body_pose=None, global_orient=None, pose2rot=True,
                return_full_pose=False, **kwargs):
        betas = self._default_betas if betas is None else betas
        body_pose = self._default_body_pose if body_pose is None else body_pose
        global_orient = self._default_global_orient if global_orient is None else global_orient
        if kwargs.get('transl') is not None:
            raise ops.TuchError('transl is not supported on the TUCH path (create_transl=False everywhere)')
        B = betas.shape[0]
        if pose2rot:
            full_pose = torch.cat([global_orient.reshape(B, -1), body_pose.reshape(B, -1)], dim=1)
        else:
            full_pose = torch.cat([global_orient.reshape(B, -1, 3, 3), body_pose.reshape(B, -1, 3, 3)], dim=1)
        verts, joints = _SmplFunction.apply(self, betas, full_pose.contiguous(), not pose2rot)
        return ModelOutput(vertices=verts, joints=joints, full_pose=full_pose if return_full_pose else None,
                           betas=betas, global_orient=global_orient, body_pose=body_pose)
