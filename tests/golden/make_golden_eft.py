"""Golden vectors for SURVEY.md 8(f) row f4 by EXECUTING the reference's own tuch/eft/loss.py
(EFTLoss.contact_loss, :129-181) on seeded synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden_eft.py   ->   tests/golden/eft_contact_loss.npz

What is real and what is stubbed: the module imports with the stub modules of make_golden.py (absent
`trimesh` / un-shipped `data` tree; tuch/eft/loss.py itself needs no third-party arithmetic).  EFTLoss.__init__
hard-codes a CUDA tensor (:49) and reads the un-shipped DSC pickles (:67-69), so the object is created with
__new__ and given exactly the attributes contact_loss reads (options.batch_size, device, face_tensor, geomask,
cdict, segments -- the latter the reference's own BatchBodySegment).  batch_pairwise_dist is CPU-patched as in
make_golden.py.  The reference hands the WHOLE batch to batch_has_self_isec inside its per-body loop (:150),
which only works for batch size 1 (the EFT fitter's setting), so every body is evaluated as its own batch of 1.
"""
import functools
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, '/root/reference')

from oracle import lbs as olbs              # noqa: E402
import make_golden as mg                    # noqa: E402


def main():
    model, geo, regions, segs, hd_reg, hd_fidx, gmm = mg.small_assets()
    cwd = os.getcwd()
    work = tempfile.mkdtemp(prefix='tuch_golden_eft_')
    tmodel = mg.install_stubs(work, model, segs, hd_reg, hd_fidx, gmm)
    import tuch.utils.contact as rc
    import tuch.utils.segmentation as rseg
    import tuch.eft.loss as rel
    rel.batch_pairwise_dist = functools.partial(rc.batch_pairwise_dist, use_cuda=False)
    g = np.load(os.path.join(HERE, 'contact_fitting_loss.npz'))
    verts = g['thres02_seg/verts']                 # the three posed bodies of the SMPLify-DC golden
    gt_contact = g['gt_contact'].copy()
    gt_contact[1] = 0                              # one body without annotated pairs (r2r term absent)
    geothres = float(g['geothres'])
    faces = torch.tensor(model['faces'])
    crit = rel.EFTLoss.__new__(rel.EFTLoss)
    torch.nn.Module.__init__(crit)
    crit.device = torch.device('cpu')
    crit.options = types.SimpleNamespace(batch_size=1)
    crit.face_tensor = faces[None]
    crit.geomask = torch.tensor(geo) > geothres
    crit.cdict = regions
    crit.segments = rseg.BatchBodySegment(list(segs.keys()), faces)
    losses, grads = [], []
    for b in range(verts.shape[0]):
        v = torch.tensor(verts[[b]], requires_grad=True)
        val = crit.contact_loss(torch.tensor(gt_contact[[b]]), v)
        val.backward()
        losses.append(val.item())
        grads.append(v.grad.numpy()[0])
    os.chdir(cwd)
    out = os.path.join(HERE, 'eft_contact_loss.npz')
    np.savez_compressed(out, verts=verts, gt_contact=gt_contact, geothres=geothres, loss=np.array(losses, np.float64),
                        g_verts=np.stack(grads).astype(np.float32))
    print('EFT contact losses', losses, 'file KB', os.path.getsize(out) // 1024)


if __name__ == '__main__':
    main()
