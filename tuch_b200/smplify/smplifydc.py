"""Host-side mirror of tuch/smplify/smplifydc.py: SMPLifyDC.__init__ (:30-66), __call__ (:68-236),
get_fitting_loss (:238-276) -- same constructor / call arguments and the same 7-tuple return.

Per iteration the reference runs ~80 ATen launches per body inside a Python loop over the batch plus
torch.optim.Adam's per-parameter loop; here one iteration is: fused LBS forward (3 launches), the
fused contact query (winding + segment whitelist + masked nearest vertex), the objective kernels
(value + analytic gradient), the fused LBS backward and one Adam kernel per parameter tensor, all
batched over the bodies and enqueued on the current stream without host synchronisation.

Where the reference reads config.SMPL_MODEL_DIR / config.PRIOR_FOLDER / constants.JOINT_IDS from its
un-shipped data tree, the same objects can be handed in directly (`smpl=`, `pose_prior=`,
`ign_joints=`).
"""
import numpy as np
import torch

from .. import ops
from ..models.smpl import SMPL
from .losses import camera_fitting_loss, body_fitting_loss, contact_fitting_loss, topology_for
from .prior import MaxMixturePrior

IGNORED_JOINT_NAMES = ['OP Neck', 'OP RHip', 'OP LHip', 'Right Hip', 'Left Hip']

# one capture stream per device: the library keeps one grow-only scratch arena per stream, so captures
# share it instead of each bringing their own
_CAPTURE_STREAMS = {}


class _Adam:
    """torch.optim.Adam(params, lr, betas) semantics over the fused Adam kernel (tuch_adam_step);
    the step counter lives on the device."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8):
        self.params = params
        self.lr, self.betas, self.eps = lr, betas, eps
        self.state = [(torch.zeros_like(p), torch.zeros_like(p),
                       torch.zeros((), dtype=torch.int32, device=p.device)) for p in params]

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def step(self):
        with torch.no_grad():
            for p, (m, v, t) in zip(self.params, self.state):
                if p.grad is None:
                    continue
                ops.adam_step(p, p.grad.contiguous(), m, v, t, self.lr, self.betas[0], self.betas[1], self.eps)


def _capture_iteration(step_eager, params, opt, warmup=2):
    """Captures one optimisation iteration (forward, loss, backward, Adam) into a CUDA graph.  The library's
    scratch arenas only grow outside captures, so `warmup` iterations run on the capture stream first;
    parameters and optimiser state are restored afterwards, i.e. capturing does not advance the optimisation."""
    saved = [p.detach().clone() for p in params]
    saved_state = [[t.clone() for t in st] for st in opt.state]
    dev = params[0].device
    s = _CAPTURE_STREAMS.get(dev)
    if s is None:
        s = _CAPTURE_STREAMS[dev] = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(max(1, warmup)):
            step_eager()
    torch.cuda.current_stream().wait_stream(s)
    with torch.no_grad():
        for p, v in zip(params, saved):
            p.copy_(v)
        for st, sv in zip(opt.state, saved_state):
            for t, v in zip(st, sv):
                t.copy_(v)
    opt.zero_grad()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        step_eager()
    return g


class CameraFit:
    """Stage 1 in flight (smplifydc.py:100-134): camera translation + shape (contact mode) or + global
    orientation, pose fixed.  Exists so that a training loop that calls SMPLifyDC every step with the same
    batch size can replay the iteration as a CUDA graph: load() swaps the batch in."""

    def __init__(self, owner, global_orient, body_pose, betas, camera_translation, init_cam_t, camera_center,
                 joints_2d, joints_conf, use_contact):
        self.owner = owner
        self.global_orient, self.body_pose, self.betas = global_orient, body_pose, betas
        self.camera_translation, self.init_cam_t, self.camera_center = camera_translation, init_cam_t, camera_center
        self.joints_2d, self.joints_conf = joints_2d, joints_conf
        camera_translation.requires_grad_(True)
        if use_contact:
            betas.requires_grad_(True)
            self.params = [betas, camera_translation]
        else:
            global_orient.requires_grad_(True)
            self.params = [global_orient, camera_translation]
        self.spw = 1.0 if use_contact else 0.0
        self.opt = _Adam(self.params, lr=owner.step_size, betas=(0.9, 0.999))
        self._graph = None

    def _step_eager(self):
        out = self.owner._forward(self.global_orient, self.body_pose, self.betas)
        loss = camera_fitting_loss(out, self.camera_translation, self.init_cam_t, self.camera_center, self.joints_2d,
                                   self.joints_conf, focal_length=self.owner.focal_length, shape_prior_weight=self.spw)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()

    def step(self):
        if self._graph is not None and self._gen != ops.scratch_generation():
            self._graph = None                      # tuch_release_scratch() freed what the graph points at
            self.capture()
        if self._graph is not None:
            self._graph.replay()
        else:
            self._step_eager()

    def capture(self, warmup=2):
        if self._graph is None:
            self._graph = _capture_iteration(self._step_eager, self.params, self.opt, warmup)
            self._gen = ops.scratch_generation()
        return self

    def load(self, init_pose, init_betas, init_cam_t, camera_center, joints_2d, joints_conf):
        with torch.no_grad():
            self.global_orient.copy_(init_pose[:, :3])
            self.body_pose.copy_(init_pose[:, 3:])
            self.betas.copy_(init_betas)
            self.camera_translation.copy_(init_cam_t)
            self.init_cam_t.copy_(init_cam_t)
            self.camera_center.copy_(camera_center)
            self.joints_2d.copy_(joints_2d)
            self.joints_conf.copy_(joints_conf)
            for st in self.opt.state:
                for t in st:
                    t.zero_()
        return self


class _NativeOptState:
    """What _capture_iteration saves and restores, for the fused iteration: the Adam state tensors of
    ops.ContactFitState in torch.optim.Adam's (exp_avg, exp_avg_sq, step) order."""

    def __init__(self, st):
        self.state = [(st.exp_avg_pose, st.exp_avg_sq_pose, st.step_pose),
                      (st.exp_avg_orient, st.exp_avg_sq_orient, st.step_orient)]

    def zero_grad(self):
        pass


class ContactFit:
    """One stage-2 optimisation in flight; see SMPLifyDC.begin_contact_fit.

    With the product's own pose prior (MaxMixturePrior or none) and contiguous fp32 parameters the whole iteration
    is ONE C-ABI call (tuch_contact_fit_step: SMPL forward, contact_fitting_loss, backward, Adam in the backward's
    last kernel -- no torch op in the loop); otherwise (a foreign pose-prior callable) the same kernels are driven
    through torch autograd, term by term.  Both give bit-identical parameters."""

    def __init__(self, owner, body_pose, global_orient, betas, camera_translation, camera_center, joints_2d,
                 joints_conf, contactlist, gt_contact, ignore_idxs, has_discrete_contact, contact_loss_weight,
                 contact_loss_return, segments, native=None):
        self.owner = owner
        self.body_pose, self.global_orient, self.betas = body_pose, global_orient, betas
        self.topo = topology_for(owner.geomask, owner.face_tensor, owner.smpl.get_num_verts(), contactlist, segments)
        tmpl = getattr(owner.smpl, 'v_template', None)
        if tmpl is not None and self.topo.F > 0 and self.topo.cluster_stats()['leaves'] == 0:
            self.topo.set_template(tmpl)           # face clusters of the hierarchical winding kernel
        can_native = (isinstance(owner.pose_prior, MaxMixturePrior) or owner.pose_prior is None) and all(
            isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
            for t in (body_pose, global_orient, betas, camera_translation, camera_center, joints_2d, joints_conf))
        self.native = can_native if native is None else (bool(native) and can_native)
        self.state = None
        if self.native:
            dev = body_pose.device
            B = body_pose.shape[0]
            self._gt_l3 = gt_contact[0] if gt_contact is not None else None
            self._has_dc, self._ignore = has_discrete_contact, ignore_idxs
            self._pair_active = None
            if self._gt_l3 is not None and has_discrete_contact is not None and len(self.topo.classes) > 0:
                self._pair_active = torch.zeros(B, len(self.topo.classes), device=dev, dtype=torch.uint8)
            self._body_active = torch.ones(B, device=dev, dtype=torch.uint8)
            self.args = dict(camera_t=camera_translation, camera_center=camera_center, joints_2d=joints_2d,
                             joints_conf=joints_conf, gt_contact=gt_contact, ignore_idxs=ignore_idxs,
                             has_discrete_contact=has_discrete_contact)
            self._refresh_masks()
            prior = owner.pose_prior._handle(dev) if owner.pose_prior is not None else None
            self.state = ops.ContactFitState(
                owner.smpl._handle(dev), self.topo, prior, body_pose.detach(), global_orient.detach(), betas.detach(),
                camera_translation.detach(), camera_center, joints_2d, joints_conf, body_active=self._body_active,
                pair_active=self._pair_active, euclthres=owner.euclthres, focal_length=owner.focal_length,
                pose_prior_weight=1.0, contact_loss_weight=contact_loss_weight,
                use_segments=segments is not None or len(self.topo.segment_names) > 0, lr=owner.step_size)
            # ContactFitState was built over these very tensors: the masks must stay the same objects
            self.opt = _NativeOptState(self.state)
            self.vertices, self.loss = self.state.vertices, self.state.loss
            self._graph = None
            return
        self.loop1_pose, self.loop1_orient = body_pose.detach().clone(), global_orient.detach().clone()
        body_pose.requires_grad_(True)
        global_orient.requires_grad_(True)
        self.opt = _Adam([body_pose, global_orient], lr=owner.step_size)
        self.args = dict(camera_t=camera_translation, camera_center=camera_center, joints_2d=joints_2d,
                         joints_conf=joints_conf, pose_prior=owner.pose_prior, cdict=contactlist,
                         gt_contact=gt_contact, ignore_idxs=ignore_idxs, has_discrete_contact=has_discrete_contact,
                         face_tensor=owner.face_tensor, focal_length=owner.focal_length,
                         contact_loss_weight=contact_loss_weight, output=contact_loss_return, segments=segments)
        self.vertices = None
        self.loss = None
        self._graph = None

    def _refresh_masks(self):
        """body_active = ~ignore_idxs (losses.py:73); pair_active = annotated pairs of the bodies that take part
        (:109-112) -- written IN PLACE into the tensors the fused iteration reads."""
        ign = self.args['ignore_idxs']
        with torch.no_grad():
            if ign is not None:
                self._body_active.copy_(~ign.bool())
            else:
                self._body_active.fill_(1)
            if self._pair_active is not None:
                act = (self.args['gt_contact'][0] == 1) & self.args['has_discrete_contact'].bool().view(-1, 1)
                self._pair_active.copy_(act & self._body_active.bool().view(-1, 1))

    def _step_eager(self):
        if self.native:
            return self.state.step()
        out = self.owner._forward(self.global_orient, self.body_pose, self.betas)
        self.vertices = out.vertices
        self.loss = contact_fitting_loss(self.body_pose, self.global_orient, self.loop1_pose, self.loop1_orient,
                                         self.betas, out.joints, self.topo, self.owner.euclthres,
                                         verts=out.vertices, **self.args)
        self.opt.zero_grad()
        self.loss.backward()
        self.opt.step()
        return self.loss

    def step(self):
        """One iteration: SMPL forward -> contact_fitting_loss -> backward -> Adam.  Returns the loss (a
        0-d device tensor; after capture() the same tensor object every time)."""
        if self._graph is not None and self._gen != ops.scratch_generation():
            self._graph = None                      # tuch_release_scratch() freed what the graph points at
            self.capture()
        if self._graph is not None:
            self._graph.replay()
            return self.loss
        return self._step_eager()

    def capture(self, warmup=2):
        """Captures one iteration into a CUDA graph: every later step() is a single graph launch (no
        per-kernel launch latency, no Python between the ~60 kernels).  The iteration has no host
        synchronisation and the library's scratch arenas only grow outside captures, so the warm-up
        iterations run on the capture stream first; parameters and optimiser state are restored
        afterwards, i.e. capture() does not advance the optimisation.  `vertices` and `loss` become static
        tensors that every replay overwrites."""
        if self._graph is None:
            self._graph = _capture_iteration(self._step_eager, [self.body_pose, self.global_orient], self.opt, warmup)
            self._gen = ops.scratch_generation()
        return self

    def load(self, init_pose, init_betas, init_cam_t, camera_center, keypoints_2d, gt_contact_l3=None,
             ignore_idxs=None, has_discrete_contact=None):
        """Re-uses this fit (and its captured graph) for a new batch of the same size: copies the inputs
        -- device tensors or pinned host tensors -- into the tensors the iteration reads and resets the
        optimiser state, like a fresh begin_contact_fit()."""
        dev = self.body_pose.device
        to = lambda t: t.to(dev, non_blocking=True)

        def put(dst, src):
            # straight into the tensor the iteration reads when no conversion is needed (one op instead of two:
            # this method is host-bound, ~25 small launches between two iterations of the end-to-end loop)
            if src.dtype == dst.dtype and src.is_contiguous() and src.shape == dst.shape:
                dst.copy_(src, non_blocking=True)
            else:
                dst.copy_(to(src))
        with torch.no_grad():
            pose = to(init_pose)
            self.global_orient.copy_(pose[:, :3])
            self.body_pose.copy_(pose[:, 3:])
            put(self.betas, init_betas)
            put(self.args['camera_t'], init_cam_t)
            put(self.args['camera_center'], camera_center)
            kp = to(keypoints_2d)
            self.args['joints_2d'].copy_(kp[:, :, :2])
            self.args['joints_conf'].copy_(kp[:, :, 2])
            if getattr(self, '_ign_index', None) is None:
                self._ign_index = torch.as_tensor(list(self.owner.ign_joints), device=dev, dtype=torch.long)
            self.args['joints_conf'].index_fill_(1, self._ign_index, 0.0)
            if gt_contact_l3 is not None:
                put(self.args['gt_contact'][0], gt_contact_l3)
            if ignore_idxs is not None:
                put(self.args['ignore_idxs'], ignore_idxs)
            if has_discrete_contact is not None:
                put(self.args['has_discrete_contact'], has_discrete_contact)
            states = [t for st in self.opt.state for t in st]
            if states:
                torch._foreach_zero_(states)
            if self.native:
                self._refresh_masks()
        return self


class SMPLifyDC():
    """SMPLify-DC optimisation follows the SMPLify routine, but takes discrete contact annotations
    into account."""

    def __init__(self,
                 step_size=1e-2,
                 batch_size=66,
                 num_iters=100,
                 focal_length=5000,
                 geodistssmpl=None,
                 geothres=0.0,
                 euclthres=0.0,
                 device=torch.device('cuda'),
                 smpl=None, pose_prior=None, ign_joints=None, use_cuda_graph=False, native_step=True):
        self.device = torch.device(device)
        # stage-2 iterations as one fused C-ABI call each (tuch_contact_fit_step); False keeps the term-by-term
        # composition through torch autograd (same kernels, bit-identical parameters)
        self.native_step = bool(native_step)
        # iterations of both stages replayed as one CUDA graph launch each (_call_graphed); pays off when the
        # iteration is launch-bound (small batches, or a training loop that fits every step)
        self.use_cuda_graph = bool(use_cuda_graph)
        self._camera_fits, self._contact_fits = {}, {}
        if self.device.type != 'cuda':
            raise ops.TuchError('SMPLifyDC needs a CUDA device: tuch_b200 has no CPU fallback')
        self.focal_length = focal_length
        self.step_size = step_size
        self.num_iters = num_iters
        if ign_joints is None or pose_prior is None or smpl is None:
            try:                                           # the reference's own sources (smplifydc.py:46-56)
                from configs import config
                from data.essentials import constants
            except Exception as e:
                raise ops.TuchError('SMPLifyDC: smpl= / pose_prior= / ign_joints= not given and the reference '
                                    'data tree (configs.config, data.essentials.constants) is not importable: %s' % (e,))
            if ign_joints is None:
                ign_joints = [constants.JOINT_IDS[i] for i in IGNORED_JOINT_NAMES]
            if pose_prior is None:
                pose_prior = MaxMixturePrior(prior_folder=config.PRIOR_FOLDER, num_gaussians=8, dtype=torch.float32)
            if smpl is None:
                smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=batch_size, create_transl=False)
        self.ign_joints = [int(i) for i in ign_joints]
        self.pose_prior = pose_prior.to(self.device)
        self.smpl = smpl.to(self.device)
        self.face_tensor = torch.tensor(self.smpl.faces.astype(np.int64), dtype=torch.long, device=self.device) \
            .unsqueeze_(0).repeat([batch_size, 1, 1])
        self.geodistssmpl = geodistssmpl
        self.geothres = geothres
        self.geomask = self.geodistssmpl > self.geothres
        self.euclthres = euclthres

    # ------------------------------------------------------------------ helpers
    def _forward(self, global_orient, body_pose, betas, **kw):
        return self.smpl(global_orient=global_orient, body_pose=body_pose, betas=betas, **kw)

    def begin_contact_fit(self, body_pose, global_orient, betas, camera_translation, camera_center, joints_2d,
                          joints_conf, contactlist, gt_contact, ignore_idxs, has_discrete_contact,
                          contact_loss_weight=1, contact_loss_return='sum', segments=None, native=None):
        """Stage-2 contact optimisation state (smplifydc.py:139-183): body_pose / global_orient become the
        optimised leaves (updated in place), everything else is held fixed.  ContactFit.step() runs one
        iteration: SMPL forward -> contact_fitting_loss -> backward -> Adam."""
        return ContactFit(self, body_pose, global_orient, betas, camera_translation, camera_center, joints_2d,
                          joints_conf, contactlist, gt_contact, ignore_idxs, has_discrete_contact,
                          contact_loss_weight, contact_loss_return, segments,
                          native=self.native_step if native is None else native)

    def __call__(self, init_pose, init_betas, init_cam_t,
                 camera_center, keypoints_2d, use_contact=False,
                 contactlist=[], gt_contact=None,
                 ignore_idxs=None, has_discrete_contact=None,
                 has_gt_keypoints=None, contact_loss_weight=1,
                 contact_loss_return='sum', segments=None):
        """Perform body fitting.  Returns (vertices, joints, pose, betas, camera_translation,
        reprojection_loss, optiverts) exactly as smplifydc.py:234."""
        if self.use_cuda_graph and use_contact and self.num_iters > 2:
            return self._call_graphed(init_pose, init_betas, init_cam_t, camera_center, keypoints_2d, contactlist,
                                      gt_contact, ignore_idxs, has_discrete_contact, has_gt_keypoints,
                                      contact_loss_weight, contact_loss_return, segments)
        camera_translation = init_cam_t.clone()
        joints_2d = keypoints_2d[:, :, :2].contiguous()
        joints_conf = keypoints_2d[:, :, -1].clone()
        body_pose = init_pose[:, 3:].detach().clone()
        global_orient = init_pose[:, :3].detach().clone()
        betas = init_betas.detach().clone()

        # ---- stage 1: camera translation + (shape | global orientation), pose fixed (:100-134)
        camera_translation.requires_grad_(True)
        if use_contact:
            betas.requires_grad_(True)
            stage1 = [betas, camera_translation]
        else:
            global_orient.requires_grad_(True)
            stage1 = [global_orient, camera_translation]
        opt = _Adam(stage1, lr=self.step_size, betas=(0.9, 0.999))
        spw = 1.0 if use_contact else 0.0
        for _ in range(self.num_iters):
            out = self._forward(global_orient, body_pose, betas)
            loss = camera_fitting_loss(out, camera_translation, init_cam_t, camera_center, joints_2d, joints_conf,
                                       focal_length=self.focal_length, shape_prior_weight=spw)
            opt.zero_grad()
            loss.backward()
            opt.step()

        # ---- stage 2 (:136-210)
        optiverts = []
        joints_conf[:, self.ign_joints] = 0.0
        if use_contact:
            camera_translation.requires_grad_(False)
            betas.requires_grad_(False)
            fit = self.begin_contact_fit(body_pose, global_orient, betas, camera_translation, camera_center,
                                         joints_2d, joints_conf, contactlist, gt_contact, ignore_idxs,
                                         has_discrete_contact, contact_loss_weight, contact_loss_return, segments)
            for _ in range(self.num_iters):
                fit.step()
                optiverts.append(fit.vertices.clone() if fit.native else fit.vertices)
        else:
            body_pose.requires_grad_(True)
            betas.requires_grad_(True)
            global_orient.requires_grad_(True)
            camera_translation.requires_grad_(False)
            opt = _Adam([body_pose, betas, global_orient], lr=self.step_size, betas=(0.9, 0.999))
            for _ in range(self.num_iters):
                out = self._forward(global_orient, body_pose, betas)
                optiverts.append(out.vertices)
                loss = body_fitting_loss(body_pose, betas, out.joints, camera_translation, camera_center,
                                         joints_2d, joints_conf, self.pose_prior, focal_length=self.focal_length)
                opt.zero_grad()
                loss.backward()
                opt.step()
        if len(optiverts) == 0:
            optiverts = None

        # ---- final reprojection score and full skin (:215-229)
        with torch.no_grad():
            out = self._forward(global_orient, body_pose, betas, return_full_pose=True)
            if has_gt_keypoints is not None:
                joints_conf[has_gt_keypoints, :25] = 0
            reprojection_loss = body_fitting_loss(body_pose, betas, out.joints, camera_translation, camera_center,
                                                  joints_2d, joints_conf, self.pose_prior,
                                                  focal_length=self.focal_length, output='reprojection')
        vertices = out.vertices.detach()
        joints = out.joints.detach()
        pose = torch.cat([global_orient, body_pose], dim=-1).detach()
        betas = betas.detach()
        return vertices, joints, pose, betas, camera_translation, reprojection_loss, optiverts

    def _call_graphed(self, init_pose, init_betas, init_cam_t, camera_center, keypoints_2d, contactlist, gt_contact,
                      ignore_idxs, has_discrete_contact, has_gt_keypoints, contact_loss_weight, contact_loss_return,
                      segments):
        """The contact branch of __call__ with both stages replayed as CUDA graphs (one launch per iteration).
        The captured iterations are kept per batch size and contact configuration, so a training loop that calls
        this every step pays for the capture once; later calls copy their batch into the tensors the graphs
        read.  Same kernels in the same order as the eager branch: identical results."""
        B = init_pose.shape[0]
        n = lambda t: None if t is None else t.detach().clone()
        joints_2d = keypoints_2d[:, :, :2].contiguous()
        joints_conf = keypoints_2d[:, :, -1].clone()
        cam = self._camera_fits.get(B)
        if cam is None:
            cam = CameraFit(self, n(init_pose[:, :3]), n(init_pose[:, 3:]), n(init_betas), n(init_cam_t), n(init_cam_t),
                            n(camera_center), joints_2d, joints_conf, use_contact=True).capture()
            self._camera_fits[B] = cam
        else:
            cam.load(init_pose, init_betas, init_cam_t, camera_center, joints_2d, joints_conf)
        for _ in range(self.num_iters):
            cam.step()
        betas = cam.betas.detach().clone()
        camera_translation = cam.camera_translation.detach().clone()

        gt_l3 = gt_contact[0] if gt_contact is not None else None
        key = (B, id(contactlist), id(segments), float(contact_loss_weight), contact_loss_return, gt_l3 is None,
               ignore_idxs is None, has_discrete_contact is None)
        hit = self._contact_fits.get(key)
        if hit is None:
            conf2 = joints_conf.clone()
            conf2[:, self.ign_joints] = 0.0
            fit = self.begin_contact_fit(n(cam.body_pose), n(cam.global_orient), betas.clone(), camera_translation.clone(),
                                         n(camera_center), joints_2d.clone(), conf2, contactlist,
                                         None if gt_contact is None else [n(gt_l3)] + list(gt_contact[1:]),
                                         n(ignore_idxs), n(has_discrete_contact), contact_loss_weight,
                                         contact_loss_return, segments).capture()
            self._contact_fits[key] = (fit, contactlist, segments)      # the keyed objects stay alive with the entry
        else:
            fit = hit[0]
            fit.load(torch.cat([cam.global_orient, cam.body_pose], dim=-1).detach(), betas, camera_translation,
                     camera_center, keypoints_2d, gt_l3, ignore_idxs, has_discrete_contact)
        optiverts = []
        for _ in range(self.num_iters):
            fit.step()
            optiverts.append(fit.vertices.clone())
        body_pose, global_orient = fit.body_pose.detach().clone(), fit.global_orient.detach().clone()
        with torch.no_grad():
            out = self._forward(global_orient, body_pose, betas, return_full_pose=True)
            conf = fit.args['joints_conf'].clone()
            if has_gt_keypoints is not None:
                conf[has_gt_keypoints, :25] = 0
            reprojection_loss = body_fitting_loss(body_pose, betas, out.joints, camera_translation, camera_center,
                                                  joints_2d, conf, self.pose_prior, focal_length=self.focal_length,
                                                  output='reprojection')
        pose = torch.cat([global_orient, body_pose], dim=-1)
        return out.vertices.detach(), out.joints.detach(), pose, betas, camera_translation, reprojection_loss, optiverts

    def get_fitting_loss(self, pose, betas, cam_t, camera_center, keypoints_2d, has_gt_keypoints=None):
        """Reprojection loss [B,49] of given body and camera parameters (smplifydc.py:238-276).  As in
        the reference, the confidences of the ignored joints are zeroed IN the caller's keypoints_2d."""
        joints_2d = keypoints_2d[:, :, :2]
        joints_conf = keypoints_2d[:, :, -1]
        joints_conf[:, self.ign_joints] = 0.
        if has_gt_keypoints is not None:
            joints_conf = joints_conf.clone()
            joints_conf[has_gt_keypoints, :25] = 0
        body_pose, global_orient = pose[:, 3:], pose[:, :3]
        with torch.no_grad():
            out = self._forward(global_orient, body_pose, betas, return_full_pose=True)
            return body_fitting_loss(body_pose, betas, out.joints, cam_t, camera_center, joints_2d, joints_conf,
                                     self.pose_prior, focal_length=self.focal_length, output='reprojection')
