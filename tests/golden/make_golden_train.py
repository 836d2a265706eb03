"""Golden vectors for the caller of the hot path by EXECUTING the reference's own train step,
tuch/train/train_module.py:112-336 (TUCH.forward_train_step), with the reference's own SMPLifyDC
(tuch/smplify/smplifydc.py), RegressorLoss (tuch/train/loss.py), FitsDict (tuch/train/fits_dict.py, real
cv2.Rodrigues), estimate_translation / perspective_projection (tuch/utils/geometry.py) and SMPL wrapper
(tuch/models/smpl.py) -- on the mixed dsc / mtp batch of tests/test_train_step_gpu.py, with and without fitting
in the loop.  Run in the build container only:

    python tests/golden/make_golden_train.py   ->   tests/golden/train_step.npz

Stubbed (absent third-party / un-shipped data), as in make_golden.py: `smplx` (LBS arithmetic from oracle.lbs, so
that piece stays unpinned), `trimesh`, `data.essentials.*`; `torchgeometry`'s two conversions come from
oracle/pose.py (the published algorithm restated; unpinned).  TUCH.__init__ reads the un-shipped DSC pickles and
datasets, so the object is created with __new__ and given the attributes forward_train_step reads; the image
regressor is the stand-in of tuch_b200.synthetic (the reference's HMR is an nn.Module argument of TUCH).  CPU
patches: batch_pairwise_dist(use_cuda=False), contact_fitting_loss(device='cpu')."""
import functools
import os
import sys
import tempfile
import types
from collections import namedtuple

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, '/root/reference')

from oracle import lbs as olbs, pose as opose         # noqa: E402
from tuch_b200 import synthetic as syn                # noqa: E402
import make_golden as mg                              # noqa: E402

B, ITERS, GEO_FIT, GEO_CRIT, EUCL = 4, 3, 0.3, 0.3, 0.02          # tests/test_train_step_gpu.py
SMPL_POSE_FLIP_PERM = []
for i in [0, 2, 1, 3, 5, 4, 6, 8, 7, 9, 11, 10, 12, 14, 13, 15, 17, 16, 19, 18, 21, 20, 23, 22]:   # public SPIN constant
    SMPL_POSE_FLIP_PERM += [3 * i, 3 * i + 1, 3 * i + 2]
Opt = namedtuple('Opt', ['batch_size', 'img_res', 'run_smplify', 'use_contact_in_the_loop',
                         'contact_in_the_loop_loss_weight', 'smplify_threshold', 'contact_loss_weight',
                         'openpose_train_weight', 'gt_train_weight', 'shape_loss_weight', 'keypoint_loss_weight',
                         'pose_loss_weight', 'beta_loss_weight'])


def main():
    model, geo, regions, segs, hd_reg, hd_fidx, gmm = mg.small_assets()
    cwd = os.getcwd()
    work = tempfile.mkdtemp(prefix='tuch_golden_train_')
    tmodel = mg.install_stubs(work, model, segs, hd_reg, hd_fidx, gmm)
    sys.modules['torchgeometry'].rotation_matrix_to_angle_axis = opose.rotation_matrix_to_angle_axis
    sys.modules['torchgeometry'].angle_axis_to_rotation_matrix = opose.angle_axis_to_rotation_matrix
    sys.modules['data.essentials.constants'].SMPL_POSE_FLIP_PERM = SMPL_POSE_FLIP_PERM
    del sys.modules['cv2']                                  # make_golden.py stubs it; FitsDict needs the real one
    import cv2                                              # noqa: F401
    import tuch.utils.contact as rc
    import tuch.smplify.losses as rl
    import tuch.smplify.smplifydc as rs
    import tuch.train.loss as rtl
    import tuch.train.train_module as rtm
    import tuch.train.fits_dict as rfd
    import tuch.models.smpl as rsmpl
    cpu_pd = functools.partial(rc.batch_pairwise_dist, use_cuda=False)
    rl.batch_pairwise_dist = cpu_pd
    rtl.batch_pairwise_dist = cpu_pd
    rtm.batch_pairwise_dist = cpu_pd
    rs.contact_fitting_loss = functools.partial(rl.contact_fitting_loss, device='cpu')
    cpu = torch.device('cpu')
    V = len(model['v_template'])
    geod = torch.tensor(geo)
    face_tensor = torch.tensor(model['faces'])[None].repeat(B, 1, 1)
    jf = lambda p, b: olbs.smpl_forward(tmodel, torch.tensor(b), torch.tensor(p[:, 3:]), torch.tensor(p[:, :3]))[1].numpy()
    batch, store = syn.make_train_batch(model, regions, B, seed=0, joints_fn=jf, img_hw=16)
    out = dict(store=store, **{'batch/' + k: np.asarray(v) for k, v in batch.items() if k != 'dataset_name'})
    for tag, run_smplify in (('fit', True), ('nofit', False)):
        o = Opt(B, 224, run_smplify, True, 2000.0, 10.5, 1.0, 0.0, 1.0, 0.5, 5.0, 1.0, 0.001)
        net = syn.make_stand_in_regressor()
        fits = rfd.FitsDict.__new__(rfd.FitsDict)
        fits.flipped_parts = torch.tensor(SMPL_POSE_FLIP_PERM, dtype=torch.int64)
        fits.fits_dict = {'dsc': torch.tensor(store.copy())}
        tuch = rtm.TUCH.__new__(rtm.TUCH)
        tuch.options, tuch.device, tuch.focal_length = o, cpu, syn.FOCAL_LENGTH
        tuch.fits_dict = fits
        tuch.modelspin = net                                  # logging only (:186-196)
        tuch.model = net
        tuch.smpl = rsmpl.SMPL('x', batch_size=B)
        tuch.geodistssmpl = geod
        tuch.smplify = rs.SMPLifyDC(step_size=1e-2, batch_size=B, num_iters=ITERS, focal_length=syn.FOCAL_LENGTH,
                                    geodistssmpl=geod, geothres=GEO_FIT, euclthres=EUCL, device=cpu)
        tuch.criterion_cospin = rtl.RegressorLoss(o, cpu, V, face_tensor, geod, geothres=GEO_CRIT, euclthres=EUCL,
                                                  face_tensor=face_tensor, use_hd=True)
        tuch.contactlists = regions
        cb = {k: (v if k == 'dataset_name' else torch.tensor(v)) for k, v in batch.items()}
        loss, losses, res = tuch.forward_train_step(cb)
        loss.backward()
        for k, v in losses.items():
            out['%s/losses/%s' % (tag, k)] = np.asarray(v.detach().reshape(-1))
        for k in ('pred_vertices', 'opt_vertices', 'pred_cam_t', 'opt_cam_t', 'gt_keypoints', 'valid_kpts_anno'):
            out['%s/out/%s' % (tag, k)] = res[k].detach().numpy()
        out[tag + '/n_optiverts'] = 0 if res['smplifyoptiverts'] is None else len(res['smplifyoptiverts'])
        out[tag + '/store'] = fits.fits_dict['dsc'].numpy()
        out[tag + '/g_weight'] = net.fc.weight.grad.numpy()
        out[tag + '/g_bias'] = net.fc.bias.grad.numpy()
        print(tag, {k: float(v.reshape(-1)[0]) for k, v in losses.items()})
    os.chdir(cwd)
    path = os.path.join(HERE, 'train_step.npz')
    np.savez_compressed(path, **out)
    print('written', path, os.path.getsize(path) // 1024, 'KB')


if __name__ == '__main__':
    main()
