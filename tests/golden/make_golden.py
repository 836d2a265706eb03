"""Generates tests/golden/*.npz by executing the REFERENCE'S OWN Python (imported read-only from
/root/reference) on seeded synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box (licence: no redistribution; /root/reference does
not exist there), so its outputs are committed as small fixtures together with this script.

What is real and what is stubbed
* REAL reference code executed: tuch/utils/contact.py, tuch/utils/geometry.py,
  tuch/smplify/losses.py, tuch/smplify/prior.py, tuch/smplify/smplifydc.py,
  tuch/utils/segmentation.py, tuch/train/loss.py, tuch/models/smpl.py (wrapper),
  tuch/train/train_module.py (contact_from_verts only).
* STUBBED (absent third-party / un-shipped data): `smplx` (SMPL forward supplied by
  oracle.lbs -- so LBS parity stays UNPINNED), `trimesh` (returns vertex colours),
  `torchgeometry`, `data.essentials.*` (synthetic assets from tuch_b200.synthetic).
* CPU patches, as SURVEY.md 8(c) documents: batch_pairwise_dist(use_cuda=False) and
  contact_fitting_loss(device='cpu') because the reference hard-codes CUDA tensors
  (contact.py:30-31, losses.py:43).
"""
import functools
import os
import pickle
import sys
import tempfile
import types
from collections import namedtuple

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from tuch_b200 import synthetic as syn          # noqa: E402
from oracle import lbs as olbs                  # noqa: E402

RINGS, SEGS = 10, 12          # V = 122, F = 240
torch.manual_seed(0)


def small_assets():
    model = syn.make_body_model(RINGS, SEGS, seed=0)
    geo = syn.make_geodesics(model['v_template'], model['faces'])
    regions = syn.make_regions(model, max_pairs=12)
    segs = syn.make_segments(model)
    hd_reg, hd_fidx = syn.make_hd_regressor(model, n_hd=300)
    gmm = syn.make_gmm()
    return model, geo, regions, segs, hd_reg, hd_fidx, gmm


def install_stubs(workdir, model, segs, hd_reg, hd_fidx, gmm):
    os.makedirs(os.path.join(workdir, 'data/essentials/spin'))
    os.makedirs(os.path.join(workdir, 'data/essentials/hd_model/smpl'))
    os.makedirs(os.path.join(workdir, 'data/models/smpl'))
    np.save(os.path.join(workdir, 'data/essentials/spin/J_regressor_extra.npy'), model['J_regressor_extra'])
    with open(os.path.join(workdir, 'data/essentials/spin/gmm_08.pkl'), 'wb') as f:
        pickle.dump(gmm, f)
    np.save(os.path.join(workdir, 'data/essentials/hd_model/smpl/smpl_neutral_hd_vert_regressor.npy'), hd_reg)
    with open(os.path.join(workdir, 'data/essentials/hd_model/smpl/smpl_neutral_hd_sample_from_mesh_out.pkl'), 'wb') as f:
        pickle.dump({'faces_vert_is_sampled_from': hd_fidx}, f)
    os.chdir(workdir)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m

    mod('data')
    mod('data.essentials')
    mod('data.essentials.constants', JOINT_NAMES=syn.JOINT_NAMES, JOINT_MAP=syn.JOINT_MAP,
        JOINT_IDS=syn.JOINT_IDS, FOCAL_LENGTH=syn.FOCAL_LENGTH, IMG_RES=syn.IMG_RES)
    sys.modules['data.essentials'].constants = sys.modules['data.essentials.constants']
    mod('data.essentials.segments')
    mod('data.essentials.segments.smpl')
    su = mod('data.essentials.segments.smpl.segm_utils',
             segments={n: dict(s['bands']) for n, s in segs.items()})
    sys.modules['data.essentials.segments.smpl'].segm_utils = su

    V = len(model['v_template'])

    def tm_load(path, process=False):
        name = os.path.basename(path)[len('smpl_segment_'):-len('.ply')]
        col = np.zeros((V, 4), np.uint8)
        col[:, 3] = 255
        col[np.asarray(segs[name]['vidx']), 0] = 255
        return types.SimpleNamespace(visual=types.SimpleNamespace(vertex_colors=col))
    mod('trimesh', load=tm_load)
    mod('torchgeometry', rotation_matrix_to_angle_axis=None, angle_axis_to_rotation_matrix=None)
    mod('cv2')

    tmodel = olbs.to_torch_model(model)
    SOut = namedtuple('SOut', ['vertices', 'joints', 'full_pose', 'betas', 'global_orient', 'body_pose'])

    class StubSMPL(torch.nn.Module):
        """smplx.SMPL stand-in: same ctor/forward surface, arithmetic from oracle.lbs."""
        def __init__(self, model_path, batch_size=1, create_transl=False, gender='neutral', **kw):
            super().__init__()
            self.faces = model['faces']
            self.batch_size = batch_size

        def get_num_verts(self):
            return V

        def forward(self, betas=None, body_pose=None, global_orient=None, pose2rot=True,
                    get_skin=True, return_full_pose=False, **kw):
            full = torch.cat([global_orient, body_pose], dim=1)
            verts, j24 = olbs.lbs(tmodel, betas, full, pose2rot=pose2rot)
            j45 = torch.cat([j24, verts[:, tmodel['extra_vertex_ids']]], dim=1)
            return SOut(verts, j45, full if return_full_pose else None, betas, global_orient, body_pose)

    def v2j(J_regressor, vertices):
        return torch.einsum('bik,ji->bjk', [vertices, J_regressor])
    mod('smplx', SMPL=StubSMPL)
    mod('smplx.lbs', vertices2joints=v2j)
    return tmodel


def main():
    model, geo, regions, segs, hd_reg, hd_fidx, gmm = small_assets()
    work = tempfile.mkdtemp(prefix='tuch_golden_')
    tmodel = install_stubs(work, model, segs, hd_reg, hd_fidx, gmm)
    V, F = len(model['v_template']), len(model['faces'])
    rng = np.random.default_rng(0)

    # ---- import the real reference modules, CPU-patched
    import tuch.utils.contact as rc
    import tuch.utils.geometry as rg
    import tuch.smplify.losses as rl
    import tuch.smplify.prior as rp
    import tuch.smplify.smplifydc as rs
    import tuch.utils.segmentation as rseg
    import tuch.train.loss as rtl
    from configs import config as rconfig
    cpu_pd = functools.partial(rc.batch_pairwise_dist, use_cuda=False)
    rl.batch_pairwise_dist = cpu_pd
    rtl.batch_pairwise_dist = cpu_pd
    rs.contact_fitting_loss = functools.partial(rl.contact_fitting_loss, device='cpu')
    cpu = torch.device('cpu')

    # ================================================================ 1. primitives
    pts = rng.normal(0, 0.4, size=(2, 37, 3)).astype(np.float32)
    tri = rng.normal(0, 0.4, size=(2, 53, 3, 3)).astype(np.float32)
    x = rng.normal(0, 1.0, size=(2, 41, 3)).astype(np.float32)
    y = rng.normal(0, 1.0, size=(2, 29, 3)).astype(np.float32)
    vt = model['v_template']
    mesh_tris = vt[model['faces']]
    q_in = (0.3 * vt[::7]).astype(np.float32)          # strictly inside (star-shaped about origin)
    q_out = (1.7 * vt[::7]).astype(np.float32)
    np.savez_compressed(
        os.path.join(OUT, 'primitives.npz'),
        pts=pts, tri=tri, x=x, y=y,
        solid_angles=rc.solid_angles(torch.tensor(pts), torch.tensor(tri)).numpy(),
        winding=rc.winding_numbers(torch.tensor(pts), torch.tensor(tri)).numpy(),
        pdist_sq=cpu_pd(torch.tensor(x), torch.tensor(y), squared=True).numpy(),
        pdist=cpu_pd(torch.tensor(x), torch.tensor(x), squared=False).numpy(),
        mesh_winding_on=rc.winding_numbers(torch.tensor(vt)[None], torch.tensor(mesh_tris)[None]).numpy()[0],
        mesh_winding_in=rc.winding_numbers(torch.tensor(q_in)[None], torch.tensor(mesh_tris)[None]).numpy()[0],
        mesh_winding_out=rc.winding_numbers(torch.tensor(q_out)[None], torch.tensor(mesh_tris)[None]).numpy()[0],
        q_in=q_in, q_out=q_out)

    # geometry.py
    rv = rng.normal(0, 0.8, size=(11, 3)).astype(np.float32)
    j3 = rng.normal(0, 0.3, size=(3, 49, 3)).astype(np.float32)
    ct = np.array([[0.1, -0.05, 45.0], [0.0, 0.1, 50.0], [-0.2, 0.0, 40.0]], np.float32)
    cc = np.full((3, 2), 112.0, np.float32)
    np.savez_compressed(
        os.path.join(OUT, 'geometry.npz'), rv=rv, j3=j3, ct=ct, cc=cc,
        rodrigues_quat=rg.batch_rodrigues(torch.tensor(rv)).numpy(),
        proj=rg.perspective_projection(torch.tensor(j3), torch.eye(3)[None].expand(3, -1, -1),
                                       torch.tensor(ct), 5000.0, torch.tensor(cc)).numpy())

    # ================================================================ 2. loss pieces
    B = 3
    inp = syn.make_smplify_inputs(
        model, regions, B, seed=3,
        joints_fn=lambda p, b: olbs.smpl_forward(tmodel, torch.tensor(b), torch.tensor(p[:, 3:]),
                                                 torch.tensor(p[:, :3]))[1].numpy())
    inp['ignore_idxs'][2] = False
    prior = rp.MaxMixturePrior(prior_folder=rconfig.PRIOR_FOLDER, num_gaussians=8, dtype=torch.float32)
    pose = torch.tensor(inp['init_pose'])
    betas = torch.tensor(inp['init_betas'])
    with torch.no_grad():
        np.savez_compressed(os.path.join(OUT, 'prior.npz'), pose=pose[:, 3:].numpy(),
                            value=prior(pose[:, 3:], betas).numpy(),
                            means=prior.means.numpy(), precisions=prior.precisions.numpy(),
                            nll_weights=prior.nll_weights.numpy())

    geod = torch.tensor(geo)
    geothres = 0.3
    geomask = geod > geothres
    face_tensor = torch.tensor(model['faces'])[None].repeat(B, 1, 1)
    segments = rseg.BatchBodySegment([k for k in segs.keys()], face_tensor[0])

    # segmentation.py on posed bodies
    verts0 = olbs.smpl_forward(tmodel, betas, pose[:, 3:], pose[:, :3])[0].detach()
    seg_ext = {}
    for b in range(B):
        for name, e in zip(segments.names, segments.batch_has_self_isec(verts0[[b]])):
            seg_ext['%s/%d' % (name, b)] = e.numpy()
    np.savez_compressed(os.path.join(OUT, 'segments.npz'), verts=verts0.numpy(),
                        **{'ext/' + k: v for k, v in seg_ext.items()},
                        **{'faces/' + n: segments.segmentation[n].segment_faces.numpy() for n in segments.names},
                        **{'vidx/' + n: segments.segmentation[n].segment_vidx for n in segments.names})

    def run_cfl(euclthres, segm, weight, ignore):
        bp = pose[:, 3:].clone().requires_grad_(True)
        go = pose[:, :3].clone().requires_grad_(True)
        out = sys.modules['tuch.models.smpl'].SMPL('x', batch_size=B)(global_orient=go, body_pose=bp, betas=betas)
        verts, joints = out.vertices, out.joints
        verts.retain_grad()
        joints.retain_grad()
        loss = rl.contact_fitting_loss(
            bp, go, bp.detach(), go.detach(), betas, joints, geomask, euclthres,
            torch.tensor(inp['init_cam_t']), torch.tensor(inp['camera_center']),
            torch.tensor(inp['keypoints_2d'][:, :, :2]), torch.tensor(inp['keypoints_2d'][:, :, 2]),
            prior, cdict=regions, gt_contact=[torch.tensor(inp['gt_contact']), None],
            ignore_idxs=torch.tensor(ignore), has_discrete_contact=torch.tensor(inp['has_discrete_contact']),
            verts=verts, face_tensor=face_tensor, device='cpu', focal_length=5000.0,
            contact_loss_weight=weight, segments=segm)
        loss.backward()
        return dict(loss=loss.item(), g_verts=verts.grad.numpy(), g_joints=joints.grad.numpy(),
                    g_body_pose=bp.grad.numpy(), g_orient=go.grad.numpy(),
                    verts=verts.detach().numpy(), joints=joints.detach().numpy())

    import tuch.models.smpl  # noqa: F401  (real wrapper over the stub smplx)
    cases = {}
    for tag, (eu, sg, w, ign) in {
        'thres02_seg': (0.02, segments, 2000.0, np.array([False, False, False])),
        'thres0_noseg': (0.0, None, 1000.0, np.array([False, False, False])),
        'thres05_seg_ignore1': (0.05, segments, 1.0, np.array([False, True, False])),
    }.items():
        r = run_cfl(eu, sg, w, ign)
        for k, v in r.items():
            cases['%s/%s' % (tag, k)] = v
    # internals recomputed with the reference primitives (same calls as losses.py:76-93)
    with torch.no_grad():
        vv = torch.tensor(cases['thres02_seg/verts'])
        ext_all, am_all, wn_all = [], [], []
        for b in range(B):
            P = cpu_pd(vv[[b]], vv[[b]], squared=True)
            wn = rc.winding_numbers(vv[[b]], vv[b][face_tensor[0]][None]).squeeze()
            P[:, ~geomask] = float('inf')
            am_all.append(torch.argmin(P, axis=1)[0].numpy())
            wn_all.append(wn.numpy())
    np.savez_compressed(os.path.join(OUT, 'contact_fitting_loss.npz'),
                        init_pose=inp['init_pose'], init_betas=inp['init_betas'],
                        init_cam_t=inp['init_cam_t'], camera_center=inp['camera_center'],
                        keypoints_2d=inp['keypoints_2d'], gt_contact=inp['gt_contact'],
                        has_discrete_contact=inp['has_discrete_contact'],
                        argmin=np.stack(am_all), winding=np.stack(wn_all), geothres=geothres,
                        **cases)

    # camera / body fitting losses
    SO = namedtuple('SO', ['joints', 'betas'])
    jt = torch.tensor(cases['thres02_seg/joints']).requires_grad_(True)
    camt = torch.tensor(inp['init_cam_t']).clone().requires_grad_(True)
    bt = betas.clone().requires_grad_(True)
    cam_est = torch.tensor(inp['init_cam_t']) + 0.3
    lc = rl.camera_fitting_loss(SO(jt, bt), camt, cam_est, torch.tensor(inp['camera_center']),
                                torch.tensor(inp['keypoints_2d'][:, :, :2]), torch.tensor(inp['keypoints_2d'][:, :, 2]),
                                focal_length=5000.0, shape_prior_weight=1.0)
    lc.backward()
    bp = pose[:, 3:].clone().requires_grad_(True)
    jt2 = torch.tensor(cases['thres02_seg/joints']).requires_grad_(True)
    bt2 = betas.clone().requires_grad_(True)
    lb = rl.body_fitting_loss(bp, bt2, jt2, torch.tensor(inp['init_cam_t']), torch.tensor(inp['camera_center']),
                              torch.tensor(inp['keypoints_2d'][:, :, :2]), torch.tensor(inp['keypoints_2d'][:, :, 2]),
                              prior, focal_length=5000.0)
    lb.backward()
    np.savez_compressed(os.path.join(OUT, 'fitting_losses.npz'), cam_est=cam_est.numpy(),
                        camera_loss=lc.item(), camera_g_joints=jt.grad.numpy(), camera_g_cam=camt.grad.numpy(),
                        camera_g_betas=bt.grad.numpy(),
                        body_loss=lb.item(), body_g_pose=bp.grad.numpy(), body_g_joints=jt2.grad.numpy(),
                        body_g_betas=bt2.grad.numpy())

    # ================================================================ 3. SMPLifyDC.__call__
    sm = {}
    for tag, use_contact, eu in (('contact', True, 0.02), ('spin', False, 0.0)):
        opt = rs.SMPLifyDC(step_size=1e-2, batch_size=B, num_iters=6, focal_length=5000.0,
                           geodistssmpl=geod, geothres=geothres, euclthres=eu, device=cpu)
        outs = opt(pose.clone(), betas.clone(), torch.tensor(inp['init_cam_t']),
                   torch.tensor(inp['camera_center']), torch.tensor(inp['keypoints_2d']),
                   use_contact=use_contact, contactlist=regions,
                   gt_contact=[torch.tensor(inp['gt_contact']), None],
                   ignore_idxs=torch.zeros(B, dtype=torch.bool),
                   has_discrete_contact=torch.tensor(inp['has_discrete_contact']),
                   has_gt_keypoints=torch.tensor([True, False, False]),
                   contact_loss_weight=2000.0, contact_loss_return='sum', segments=segments)
        for n, t in zip(['vertices', 'joints', 'pose', 'betas', 'cam_t', 'reproj'], outs[:6]):
            sm['%s/%s' % (tag, n)] = t.detach().numpy()
        sm['%s/n_optiverts' % tag] = len(outs[6])
        kp = torch.tensor(inp['keypoints_2d']).clone()
        sm['%s/get_fitting_loss' % tag] = opt.get_fitting_loss(
            pose.clone(), betas.clone(), torch.tensor(inp['init_cam_t']), torch.tensor(inp['camera_center']),
            kp, has_gt_keypoints=torch.tensor([True, False, False])).numpy()
        sm['%s/kp_after' % tag] = kp.numpy()
        sm['ign_joints'] = np.array(opt.ign_joints)
    np.savez_compressed(os.path.join(OUT, 'smplify_dc.npz'), **sm)

    # ================================================================ 4. RegressorLoss.contact_loss
    Opt = namedtuple('Opt', ['contact_loss_weight'])
    rgl = {}
    for tag, use_hd in (('hd', True), ('nohd', False)):
        crit = rtl.RegressorLoss(Opt(1.0), cpu, V, face_tensor, geod, geothres=geothres,
                                 euclthres=0.02, face_tensor=face_tensor, use_hd=use_hd)
        pv = verts0.clone().requires_grad_(True)
        valid = torch.tensor([True, False, True])
        try:
            val = crit.contact_loss(pv, valid)
            val.backward()
            rgl[tag + '/loss'] = val.item()
            rgl[tag + '/g_verts'] = pv.grad.numpy()
        except Exception as e:      # the non-HD branch of the reference is broken (loss.py:303 dim=2 on 2-D)
            rgl[tag + '/error'] = np.array(repr(e))
        rgl['valid'] = valid.numpy()
    np.savez_compressed(os.path.join(OUT, 'regressor_contact_loss.npz'), verts=verts0.numpy(), **rgl)

    # contact_from_verts (train_module.py:69-91) without constructing TUCH
    import tuch.train.train_module as rtm
    fake = types.SimpleNamespace(contactlists=regions, device=cpu)
    rtm.batch_pairwise_dist = cpu_pd
    cfv = rtm.TUCH.contact_from_verts(fake, verts0)
    np.savez_compressed(os.path.join(OUT, 'contact_from_verts.npz'), verts=verts0.numpy(), value=cfv.numpy())

    # the small asset set itself is regenerated deterministically by the tests; store a digest
    np.savez_compressed(os.path.join(OUT, 'assets_digest.npz'), rings=RINGS, segs=SEGS,
                        v_sum=np.float64(model['v_template'].astype(np.float64).sum()),
                        geo_sum=np.float64(geo.astype(np.float64).sum()),
                        n_classes=len(regions['classes']))
    print('golden vectors written to', OUT)
    for f in sorted(os.listdir(OUT)):
        if f.endswith('.npz'):
            print('  %-34s %8d B' % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == '__main__':
    main()
