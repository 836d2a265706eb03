"""Eager stage-1 (camera + shape) iterations at a given batch, for an ncu launch list."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                                        # noqa: E402
from tuch_b200.smplify.smplifydc import CameraFit   # noqa: E402


class A:
    gpus, steps, warmup, batch = 1, 10, 3, 256


rig = bench.Rig(A)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
a = bench.make_assets(B, seed=1000)
s = rig.stack(a, B, num_iters=10)
d = {k: rig.t(a['inp'][k]) for k in bench.INPUT_KEYS}
kp = d['keypoints_2d']
n = lambda t: t.detach().clone()
cam = CameraFit(s['smplify'], n(d['init_pose'][:, :3]), n(d['init_pose'][:, 3:]), n(d['init_betas']), n(d['init_cam_t']),
                n(d['init_cam_t']), n(d['camera_center']), kp[:, :, :2].contiguous(), kp[:, :, 2].clone(), use_contact=True)
ms, _ = rig.timed(cam.step, 10, warmup=3)
print('stage-1 eager B=%d: %.3f ms' % (B, ms))
