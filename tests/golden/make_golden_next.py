"""Records golden vectors for the SURVEY.md 8(f) "next" rows by EXECUTING the reference's own code.

* estimate_translation: the real tuch/utils/geometry.py:156-205 (imports cleanly).
* FitsDict.rotate_pose / flip_pose: the real tuch/train/fits_dict.py:89-119, with the absent third-party
  `torchgeometry` stubbed by oracle/pose.py's restatement and the real cv2.Rodrigues.
Run in the build container (needs /root/reference): python tests/golden/make_golden_next.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

from oracle import pose as op                      # noqa: E402
from tuch_b200 import synthetic as syn             # noqa: E402

SMPL_POSE_FLIP_PERM = []
for i in [0, 2, 1, 3, 5, 4, 6, 8, 7, 9, 11, 10, 12, 14, 13, 15, 17, 16, 19, 18, 21, 20, 23, 22]:   # public SPIN constant
    SMPL_POSE_FLIP_PERM += [3 * i, 3 * i + 1, 3 * i + 2]


def main():
    rng = np.random.default_rng(42)
    # ---- estimate_translation
    import tuch.utils.geometry as rg
    B = 12
    S = torch.tensor(rng.normal(0, 0.4, size=(B, 49, 3)).astype(np.float32))
    t_true = np.tile(np.array([[0.05, -0.1, 40.0]], np.float32), (B, 1)) + rng.normal(0, 0.5, size=(B, 3)).astype(np.float32)
    p = S.numpy() + t_true[:, None]
    kp = np.zeros((B, 49, 3), np.float32)
    kp[:, :, :2] = 5000.0 * p[:, :, :2] / p[:, :, 2:3] + 112.0 + rng.normal(0, 1.0, size=(B, 49, 2))
    kp[:, :, 2] = rng.uniform(0, 1, size=(B, 49))
    kp[3, :25, 2] = 0.0                       # no OpenPose confidence at all -> zeros when not annotated
    kp[5, 25:, 2] = 0.0
    has = np.array([1, 0, 1, 0, 0, 1, 1, 0, 1, 0, 0, 1], bool)
    et = rg.estimate_translation(S, torch.tensor(kp), focal_length=5000., img_size=224., has_2d_kp_anno=torch.tensor(has))
    # ---- FitsDict pose transforms
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m
    mod('torchgeometry', angle_axis_to_rotation_matrix=op.angle_axis_to_rotation_matrix,
        rotation_matrix_to_angle_axis=op.rotation_matrix_to_angle_axis)
    mod('data')
    mod('data.essentials')
    mod('data.essentials.constants', SMPL_POSE_FLIP_PERM=SMPL_POSE_FLIP_PERM)
    sys.modules['data.essentials'].constants = sys.modules['data.essentials.constants']
    from tuch.train.fits_dict import FitsDict
    fd = FitsDict.__new__(FitsDict)
    fd.flipped_parts = torch.tensor(SMPL_POSE_FLIP_PERM, dtype=torch.int64)
    N = 16
    pose = torch.tensor(syn.fold_arms_pose(N, seed=9))
    pose[0, :3] = 0.0                                          # identity orientation (Taylor branch)
    pose[1, :3] = torch.tensor([0.0, 0.0, np.pi - 1e-4])       # near pi
    pose[2, :3] = torch.tensor([2.2, -2.2, 0.1])
    rot = torch.tensor(rng.uniform(-60, 60, size=N).astype(np.float32))
    rot[4] = 0.0
    flipped = torch.tensor(rng.integers(0, 2, size=N).astype(np.uint8))
    got = fd.flip_pose(fd.rotate_pose(pose.clone(), rot), flipped)                 # __getitem__ order
    back = fd.rotate_pose(fd.flip_pose(got.clone(), flipped), -rot)                # __setitem__ order
    np.savez_compressed(os.path.join(HERE, 'pose_bookkeeping.npz'), S=S.numpy(), kp=kp, has=has, et=et.numpy(),
                        flip_perm=np.asarray(SMPL_POSE_FLIP_PERM, np.int32), pose=pose.numpy(), rot=rot.numpy(),
                        flipped=flipped.numpy(), got=got.numpy(), back=back.numpy())
    print('estimate_translation', et[:3], '\nfits', got[:2, :6], (back - pose).abs().max())


if __name__ == '__main__':
    main()
