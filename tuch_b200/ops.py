"""Thin torch-tensor front end of the C ABI (include/tuch_b200.h).

torch is plumbing only: it owns device memory and the current stream; every computation below
is a hand-written sm_100a kernel inside libtuch_b200.so.  All functions require CUDA tensors and
raise TuchError otherwise -- there is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import lib, check, TuchError, ContactFitArgs


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TuchError('%s must be a CUDA tensor: tuch_b200 has no CPU fallback' % name)
    return t


def _f32(t, name):
    t = _dev(t, name).detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _i32_host(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.int32)


def _hp(a):
    return a.ctypes.data_as(C.c_void_p)


def device_info():
    sm, mj, mn = C.c_int(), C.c_int(), C.c_int()
    check(lib().tuch_device_info(C.byref(sm), C.byref(mj), C.byref(mn)), 'tuch_device_info')
    return dict(sm_count=sm.value, cc=(mj.value, mn.value))


def launch_count():
    return int(lib().tuch_launch_count())


def release_scratch():
    """Frees the library's scratch arenas on the current device.  Graphs captured before this call are invalid
    afterwards (scratch_generation() changes; ContactFit / CameraFit re-capture on their next step)."""
    check(lib().tuch_release_scratch(), 'tuch_release_scratch')


def scratch_generation():
    return int(lib().tuch_scratch_generation())


def kernel_timing(enable=None, reset=False):
    """Per-kernel CUDA-event timing inside the library (bench.py's roofline leg)."""
    if reset:
        check(lib().tuch_kernel_timing_reset(), 'tuch_kernel_timing_reset')
    if enable is not None:
        check(lib().tuch_kernel_timing_enable(int(bool(enable))), 'tuch_kernel_timing_enable')


def kernel_times():
    """-> {name: (total device ms, launches)} of everything timed since the last reset."""
    buf = C.create_string_buffer(4096)
    check(lib().tuch_kernel_timing_names(buf, len(buf)), 'tuch_kernel_timing_names')
    names = [n for n in buf.value.decode().split(',') if n]
    return {n: kernel_time(n) for n in names}


def kernel_time(name):
    """-> (total device ms, launches) of a timed kernel since the last reset."""
    ms, n = C.c_double(), C.c_longlong()
    check(lib().tuch_kernel_timing_read(name.encode(), C.byref(ms), C.byref(n)), 'tuch_kernel_timing_read')
    return ms.value, n.value


def strip_stream(faces):
    """Host-only: the triangle-strip vertex stream the library builds for a face list ->
    (vid int32 [L], flag uint32 [L], n_strips).  See tuch_strip_stream_host."""
    f = _i32_host(np.asarray(faces).reshape(-1, 3))
    L, n = C.c_int(), C.c_int()
    check(lib().tuch_strip_stream_host(_hp(f), len(f), None, None, 0, C.byref(L), C.byref(n)), 'tuch_strip_stream_host')
    vid = np.empty(L.value, np.int32)
    flag = np.empty(L.value, np.uint32)
    check(lib().tuch_strip_stream_host(_hp(f), len(f), _hp(vid), _hp(flag), L.value, C.byref(L), C.byref(n)),
          'tuch_strip_stream_host')
    return vid, flag, n.value


# ------------------------------------------------------------------ a1-a3 forward kernels
def cluster_tree(faces, verts):
    """Host-only: the hierarchy of clusters.cu -> dict(leaf_face[K,16], mid_off[NM+1], top_off[NT+1], vtile[T,32])."""
    f = _i32_host(np.asarray(faces).reshape(-1, 3))
    v = np.ascontiguousarray(np.asarray(verts).reshape(-1, 3), dtype=np.float32)
    k, nm, nt, t = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
    refs = (C.byref(k), C.byref(nm), C.byref(nt), C.byref(t))
    check(lib().tuch_cluster_tree_host(_hp(f), len(f), len(v), _hp(v), None, 0, None, 0, None, 0, None, 0, *refs),
          'tuch_cluster_tree_host')
    leaf = np.empty((k.value, 16), np.int32)
    mid = np.empty(nm.value + 1, np.int32)
    top = np.empty(nt.value + 1, np.int32)
    vt = np.empty((t.value, 32), np.int32)
    check(lib().tuch_cluster_tree_host(_hp(f), len(f), len(v), _hp(v), _hp(leaf), k.value, _hp(mid), nm.value, _hp(top),
                                       nt.value, _hp(vt), t.value, *refs), 'tuch_cluster_tree_host')
    return dict(leaf_face=leaf, mid_off=mid, top_off=top, vtile=vt)


def pairwise_dist(x, y, squared=True):
    x, y = _f32(x, 'x'), _f32(y, 'y')
    if x.dim() != 3 or y.dim() != 3 or x.shape[2] != 3 or y.shape[2] != 3 or x.shape[0] != y.shape[0]:
        raise TuchError('pairwise_dist expects x[bs,Nx,3], y[bs,Ny,3]; got %s, %s' % (tuple(x.shape), tuple(y.shape)))
    bs, nx, ny = x.shape[0], x.shape[1], y.shape[1]
    P = torch.empty(bs, nx, ny, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(lib().tuch_pairwise_dist(_ptr(x), _ptr(y), bs, nx, ny, int(bool(squared)), _ptr(P), _stream()),
              'tuch_pairwise_dist')
    return P


def pairwise_dist_backward(x, y, P, gP, squared=True):
    x, y, gP = _f32(x, 'x'), _f32(y, 'y'), _f32(gP, 'gP')
    P = _f32(P, 'P') if P is not None else None             # only the sqrt form reads it
    if P is None and not squared:
        raise TuchError('pairwise_dist_backward: the sqrt form needs P')
    bs, nx, ny = x.shape[0], x.shape[1], y.shape[1]
    gx, gy = torch.empty_like(x), torch.empty_like(y)
    with torch.cuda.device(x.device):
        check(lib().tuch_pairwise_dist_backward(_ptr(x), _ptr(y), _ptr(P), _ptr(gP), bs, nx, ny,
                                                int(bool(squared)), _ptr(gx), _ptr(gy), _stream()),
              'tuch_pairwise_dist_backward')
    return gx, gy


def solid_angles(points, triangles):
    p, t = _f32(points, 'points'), _f32(triangles, 'triangles')
    if p.dim() != 3 or t.dim() != 4 or p.shape[2] != 3 or tuple(t.shape[2:]) != (3, 3) or p.shape[0] != t.shape[0]:
        raise TuchError('solid_angles expects points[B,Q,3], triangles[B,F,3,3]; got %s, %s'
                        % (tuple(p.shape), tuple(t.shape)))
    B, Q, F = p.shape[0], p.shape[1], t.shape[1]
    out = torch.empty(B, Q, F, device=p.device, dtype=torch.float32)
    with torch.cuda.device(p.device):
        check(lib().tuch_solid_angles(_ptr(p), _ptr(t), B, Q, F, _ptr(out), _stream()), 'tuch_solid_angles')
    return out


def winding_numbers(points, triangles):
    p, t = _f32(points, 'points'), _f32(triangles, 'triangles')
    if p.dim() != 3 or t.dim() != 4 or p.shape[2] != 3 or tuple(t.shape[2:]) != (3, 3) or p.shape[0] != t.shape[0]:
        raise TuchError('winding_numbers expects points[B,Q,3], triangles[B,F,3,3]; got %s, %s'
                        % (tuple(p.shape), tuple(t.shape)))
    B, Q, F = p.shape[0], p.shape[1], t.shape[1]
    out = torch.empty(B, Q, device=p.device, dtype=torch.float32)
    with torch.cuda.device(p.device):
        check(lib().tuch_winding_numbers(_ptr(p), _ptr(t), B, Q, F, _ptr(out), _stream()), 'tuch_winding_numbers')
    return out


# ------------------------------------------------------------------ topology handle
class Topology:
    """Device-resident constants of one mesh topology (faces, geodesic mask, DSC regions, body
    segments) -- the arguments the reference threads through every call as face_tensor / geomask /
    cdict / segments (smplifydc.py:58-66, losses.py:34-51)."""

    def __init__(self, faces, num_verts, device):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise TuchError('Topology needs a CUDA device: tuch_b200 has no CPU fallback')
        f = faces.detach().cpu().numpy() if isinstance(faces, torch.Tensor) else np.asarray(faces)
        f = _i32_host(f.reshape(-1, 3))
        self.V, self.F = int(num_verts), int(len(f))
        self.faces_np = f
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().tuch_topology_create(self.V, self.F, _hp(f), C.byref(self._h)), 'tuch_topology_create')
        self.has_mask = False
        self.region_names = []
        self.classes = []
        self.segment_names = []
        self.segment_vidx = []
        self.n_seg_verts = 0
        self.n_hd = 0

    def __del__(self):
        try:
            h = getattr(self, '_h', None)
            if h is not None and h.value:
                lib().tuch_topology_destroy(h)
                self._h = None
        except Exception:       # interpreter shutdown: modules may already be torn down
            pass

    @property
    def handle(self):
        return self._h

    # winding-number evaluation mode of contact_query (include/tuch_b200.h)
    WINDING_EXACT, WINDING_FAST = 0, 1

    def set_template(self, verts):
        """Clusters the faces on these [V,3] positions (the model's v_template) for the hierarchical
        winding kernel; without it the first queried body is used."""
        v = verts.detach().cpu().numpy() if isinstance(verts, torch.Tensor) else np.asarray(verts)
        v = np.ascontiguousarray(v.reshape(-1, 3), dtype=np.float32)
        if v.shape != (self.V, 3):
            raise TuchError('template must be [%d,3], got %s' % (self.V, tuple(v.shape)))
        with torch.cuda.device(self.device):
            check(lib().tuch_topology_set_template(self._h, _hp(v)), 'tuch_topology_set_template')

    def set_winding_mode(self, mode):
        check(lib().tuch_topology_set_winding_mode(self._h, int(mode)), 'tuch_topology_set_winding_mode')

    def cluster_stats(self):
        k, nm, nt, t, lf = (C.c_int32(0) for _ in range(5))
        check(lib().tuch_topology_cluster_stats(self._h, C.byref(k), C.byref(nm), C.byref(nt), C.byref(t), C.byref(lf)),
              'tuch_topology_cluster_stats')
        return dict(leaves=int(k.value), mids=int(nm.value), tops=int(nt.value), vertex_tiles=int(t.value),
                    leaf_faces=int(lf.value))

    def query_stats(self):
        """-> dict(refine_vertices, refine_points): queries the last hierarchical call re-evaluated exactly."""
        a, b = C.c_int32(0), C.c_int32(0)
        with torch.cuda.device(self.device):
            check(lib().tuch_topology_query_stats(self._h, C.byref(a), C.byref(b), _stream()), 'tuch_topology_query_stats')
        return dict(refine_vertices=int(a.value), refine_points=int(b.value))

    def pack_nodes(self, verts, direct=False):
        """-> [B, tops + mids + leaves, 28] node records of the face hierarchy for verts [B,V,3] (test hook)."""
        v = _f32(verts, 'verts')
        st = self.cluster_stats()
        out = torch.empty(v.shape[0], st['tops'] + st['mids'] + st['leaves'], 28, device=v.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(lib().tuch_topology_pack_nodes(self._h, _ptr(v), int(v.shape[0]), int(bool(direct)), _ptr(out), _stream()),
                  'tuch_topology_pack_nodes')
        return out

    # geomask = geodist > geothres (smplifydc.py:65)
    def set_geodist(self, geodist, geothres):
        g = _f32(geodist.to(self.device) if isinstance(geodist, torch.Tensor) else
                 torch.as_tensor(np.asarray(geodist), device=self.device), 'geodist')
        if tuple(g.shape) != (self.V, self.V):
            raise TuchError('geodist must be [%d,%d], got %s' % (self.V, self.V, tuple(g.shape)))
        with torch.cuda.device(self.device):
            check(lib().tuch_topology_set_geodist(self._h, _ptr(g), float(geothres), _stream()),
                  'tuch_topology_set_geodist')
            torch.cuda.current_stream().synchronize()
        self.has_mask = True

    def set_geomask(self, geomask):
        m = geomask if isinstance(geomask, torch.Tensor) else torch.as_tensor(np.asarray(geomask))
        m = _dev(m.to(self.device), 'geomask').to(torch.uint8).contiguous()
        if tuple(m.shape) != (self.V, self.V):
            raise TuchError('geomask must be [%d,%d], got %s' % (self.V, self.V, tuple(m.shape)))
        with torch.cuda.device(self.device):
            check(lib().tuch_topology_set_geomask(self._h, _ptr(m), _stream()), 'tuch_topology_set_geomask')
            torch.cuda.current_stream().synchronize()
        self.has_mask = True

    def set_regions(self, cdict):
        """cdict = {'classes': sequence of (regionA, regionB), 'csig': {region: vertex ids}}
        (losses.py:110-113, train_module.py:65-67)."""
        names = []
        for ra, rb in cdict['classes']:
            for r in (ra, rb):
                if r not in names:
                    names.append(r)
        off, ids = [0], []
        for n in names:
            v = [int(i) for i in cdict['csig'][n]]
            ids += v
            off.append(len(ids))
        idx = {n: i for i, n in enumerate(names)}
        pa = _i32_host([idx[a] for a, _ in cdict['classes']])
        pb = _i32_host([idx[b] for _, b in cdict['classes']])
        off, ids = _i32_host(off), _i32_host(ids)
        with torch.cuda.device(self.device):
            check(lib().tuch_topology_set_regions(self._h, len(names), _hp(off), _hp(ids), len(pa), _hp(pa), _hp(pb)),
                  'tuch_topology_set_regions')
        self.region_names = names
        self.classes = [tuple(c) for c in cdict['classes']]

    def set_segments(self, segments):
        """segments: iterable of (name, vidx, closed_faces[n,3], band loops) in BatchBodySegment order
        (segmentation.py:113-115); closed_faces index V + k for the k-th band centroid."""
        v_off, v_ids, f_off, f_ids, b_off, l_off, l_ids = [0], [], [0], [], [0], [0], []
        names, vidx_list = [], []
        for name, vidx, cfaces, loops in segments:
            names.append(name)
            vidx = [int(i) for i in vidx]
            vidx_list.append(np.asarray(vidx, dtype=np.int64))
            v_ids += vidx
            v_off.append(len(v_ids))
            cf = np.asarray(cfaces, dtype=np.int64).reshape(-1, 3)
            f_ids += [int(i) for i in cf.reshape(-1)]
            f_off.append(len(f_ids) // 3)
            for lp in loops:
                l_ids += [int(i) for i in lp]
                l_off.append(len(l_ids))
            b_off.append(len(l_off) - 1)
        arrs = [_i32_host(a) for a in (v_off, v_ids, f_off, f_ids, b_off, l_off, l_ids)]
        with torch.cuda.device(self.device):
            check(lib().tuch_topology_set_segments(self._h, len(names), *[_hp(a) for a in arrs]),
                  'tuch_topology_set_segments')
        self.segment_names = names
        self.segment_vidx = vidx_list
        self.n_seg_verts = int(v_off[-1])

    def set_hd(self, regressor, faces_vert_is_sampled_from):
        """HD-point model of the regressor loss (loss.py:81-89): `regressor` is the dense [N_hd, V] matrix
        the reference loads (numpy / torch / scipy.sparse); only its non-zeros are kept (CSR)."""
        if hasattr(regressor, 'tocsr'):
            csr = regressor.tocsr()
            off, cols, vals = csr.indptr, csr.indices, csr.data
            n = csr.shape[0]
        else:
            R = regressor.detach().cpu().numpy() if isinstance(regressor, torch.Tensor) else np.asarray(regressor)
            if R.ndim != 2 or R.shape[1] != self.V:
                raise TuchError('HD regressor must be [N_hd,%d], got %s' % (self.V, R.shape))
            n = R.shape[0]
            r, c = np.nonzero(R)
            off = np.zeros(n + 1, np.int64)
            np.add.at(off, r + 1, 1)
            off = np.cumsum(off)
            cols, vals = c, R[r, c]
        hf = faces_vert_is_sampled_from
        hf = hf.detach().cpu().numpy() if isinstance(hf, torch.Tensor) else np.asarray(hf)
        if len(hf) != n:
            raise TuchError('faces_vert_is_sampled_from has %d entries for %d HD points' % (len(hf), n))
        off, cols, hf = _i32_host(off), _i32_host(cols), _i32_host(hf)
        vals = np.ascontiguousarray(vals, dtype=np.float32)
        with torch.cuda.device(self.device):
            check(lib().tuch_topology_set_hd(self._h, int(n), _hp(off), _hp(cols), _hp(vals), _hp(hf)),
                  'tuch_topology_set_hd')
        self.n_hd = int(n)

    def regressor_contact_loss(self, verts, valid=None, euclthres=0.02, use_hd=True, weight=1.0, g_loss=None,
                               g_verts=None, debug=False):
        """loss.py:240-315 for the whole batch -> loss[B] (0 for invalid bodies); see tuch_regressor_contact_loss."""
        v = self._verts(verts)
        B = v.shape[0]
        val = _dev(valid, 'valid').to(torch.uint8).contiguous() if valid is not None else None
        gl = _f32(g_loss, 'g_loss') if g_loss is not None else None
        loss = torch.zeros(B, device=v.device, dtype=torch.float32)
        N = getattr(self, 'n_hd', 0) if use_hd else 0
        dbg = {}
        if debug and use_hd:
            dbg = dict(counts=torch.zeros(B, device=v.device, dtype=torch.int32),
                       sel=torch.zeros(B, N, device=v.device, dtype=torch.int32),
                       hd_argmin=torch.zeros(B, N, device=v.device, dtype=torch.int32),
                       hd_exterior=torch.zeros(B, N, device=v.device, dtype=torch.uint8))
        with torch.cuda.device(v.device):
            check(lib().tuch_regressor_contact_loss(self._h, _ptr(v), B, _ptr(val), float(euclthres), int(bool(use_hd)),
                                                    float(weight), _ptr(gl), _ptr(loss), _ptr(g_verts),
                                                    _ptr(dbg.get('counts')), _ptr(dbg.get('sel')),
                                                    _ptr(dbg.get('hd_argmin')), _ptr(dbg.get('hd_exterior')), _stream()),
                  'tuch_regressor_contact_loss')
        return (loss, dbg) if debug else loss

    # ------------------------------------------------------------------ queries
    def _verts(self, verts):
        v = _f32(verts, 'verts')
        if v.dim() != 3 or v.shape[1] != self.V or v.shape[2] != 3:
            raise TuchError('verts must be [B,%d,3], got %s' % (self.V, tuple(v.shape)))
        if v.device != self.device:
            raise TuchError('verts live on %s but the topology on %s' % (v.device, self.device))
        return v

    def contact_query(self, verts, use_segments=True, want_nearest=True, want_winding=True, within=None):
        """Fused losses.py:76-93 for the whole batch -> dict(argmin int32[B,V], min_sq[B,V],
        winding[B,V], exterior bool[B,V]).  within=r (metres): the nearest vertex only where losses.py:96-103
        consumes it -- interior vertices and vertices with an allowed vertex within r; (-1, inf) elsewhere."""
        v = self._verts(verts)
        B = v.shape[0]
        out = {}
        am = mn = w = ext = None
        if want_nearest:
            am = torch.empty(B, self.V, device=v.device, dtype=torch.int32)
            mn = torch.empty(B, self.V, device=v.device, dtype=torch.float32)
        if want_winding:
            w = torch.empty(B, self.V, device=v.device, dtype=torch.float32)
            ext = torch.empty(B, self.V, device=v.device, dtype=torch.uint8)
        with torch.cuda.device(v.device):
            if within is not None:
                if not (want_nearest and want_winding):
                    raise TuchError('contact_query(within=...) needs the nearest-vertex and the winding outputs')
                check(lib().tuch_contact_query_within(self._h, _ptr(v), B, int(bool(use_segments)), float(within), _ptr(am),
                                                      _ptr(mn), _ptr(w), _ptr(ext), _stream()), 'tuch_contact_query_within')
            else:
                check(lib().tuch_contact_query(self._h, _ptr(v), B, int(bool(use_segments)), _ptr(am), _ptr(mn),
                                               _ptr(w), _ptr(ext), _stream()), 'tuch_contact_query')
        out.update(argmin=am, min_sq=mn, winding=w, exterior=None if ext is None else ext.bool())
        return out

    def segment_exterior(self, verts):
        """BatchBodySegment.batch_has_self_isec for a batch: list (segment order) of bool [B, n_s]."""
        v = self._verts(verts)
        B = v.shape[0]
        flags = torch.empty(B, self.n_seg_verts, device=v.device, dtype=torch.uint8)
        wind = torch.empty(B, self.n_seg_verts, device=v.device, dtype=torch.float32)
        with torch.cuda.device(v.device):
            check(lib().tuch_segment_exterior(self._h, _ptr(v), B, _ptr(flags), _ptr(wind), _stream()),
                  'tuch_segment_exterior')
        sizes = [len(x) for x in self.segment_vidx]
        return [f.bool() for f in torch.split(flags, sizes, dim=1)], list(torch.split(wind, sizes, dim=1))

    def region_min(self, verts, masked=True, active=None):
        """-> (min_sq[B,n_pairs], arg_i, arg_j) over the annotated region pairs."""
        v = self._verts(verts)
        B, P = v.shape[0], len(self.classes)
        mn = torch.zeros(B, P, device=v.device, dtype=torch.float32)
        ai = torch.full((B, P), -1, device=v.device, dtype=torch.int32)
        aj = torch.full((B, P), -1, device=v.device, dtype=torch.int32)
        act = None
        if active is not None:
            act = _dev(active, 'active').to(torch.uint8).contiguous()
            if tuple(act.shape) != (B, P):
                raise TuchError('active must be [%d,%d], got %s' % (B, P, tuple(act.shape)))
        with torch.cuda.device(v.device):
            check(lib().tuch_region_min(self._h, _ptr(v), B, int(bool(masked)), _ptr(act), _ptr(mn), _ptr(ai),
                                        _ptr(aj), _stream()), 'tuch_region_min')
        return mn, ai, aj


# ------------------------------------------------------------------ SMPL body model handle
class SmplHandle:
    """Device-resident SMPL model (tuch_smpl): constants uploaded once, forward / backward kernels."""

    def __init__(self, arrays, device):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise TuchError('SmplHandle needs a CUDA device: tuch_b200 has no CPU fallback')
        f32 = lambda a: np.ascontiguousarray(np.asarray(a), dtype=np.float32)
        vt = f32(arrays['v_template'])
        self.V = int(vt.shape[0])
        sd = f32(arrays['shapedirs'])
        self.L = int(sd.shape[-1])
        pd = f32(arrays['posedirs']).reshape(207, self.V * 3)
        jr = f32(arrays['J_regressor'])
        lw = f32(arrays['lbs_weights'])
        par = _i32_host(arrays['parents'])
        xv = _i32_host(arrays['extra_vertex_ids'])
        xr = f32(arrays['J_regressor_extra']).reshape(-1, self.V)
        jm = _i32_host(arrays['joint_map'])
        self.NO = int(len(jm))
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().tuch_smpl_create(self.V, self.L, _hp(vt), _hp(sd), _hp(pd), _hp(jr), _hp(lw), _hp(par),
                                         len(xv), _hp(xv), xr.shape[0], _hp(xr), len(jm), _hp(jm),
                                         C.byref(self._h)), 'tuch_smpl_create')

    def __del__(self):
        try:
            h = getattr(self, '_h', None)
            if h is not None and h.value:
                lib().tuch_smpl_destroy(h)
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def workspace(self, B):
        n = int(lib().tuch_smpl_workspace_floats(self._h, int(B)))
        return torch.empty(max(n, 1), device=self.device, dtype=torch.float32)

    def forward(self, betas, pose, is_rotmat, workspace=None):
        """betas[B,L], pose [B,72] | [B,24,3,3] -> (vertices[B,V,3], joints[B,NO,3], workspace)."""
        betas, pose = _f32(betas, 'betas'), _f32(pose, 'pose')
        B = betas.shape[0]
        if betas.shape[1] != self.L:
            raise TuchError('betas must be [B,%d], got %s' % (self.L, tuple(betas.shape)))
        if pose.numel() != B * (216 if is_rotmat else 72):
            raise TuchError('pose has %d elements for batch %d (is_rotmat=%s)' % (pose.numel(), B, is_rotmat))
        ws = workspace if workspace is not None else self.workspace(B)
        verts = torch.empty(B, self.V, 3, device=betas.device, dtype=torch.float32)
        joints = torch.empty(B, self.NO, 3, device=betas.device, dtype=torch.float32)
        with torch.cuda.device(betas.device):
            check(lib().tuch_smpl_forward(self._h, _ptr(betas), _ptr(pose), int(bool(is_rotmat)), B, _ptr(ws),
                                          _ptr(verts), _ptr(joints), _stream()), 'tuch_smpl_forward')
        return verts, joints, ws

    def backward(self, pose, is_rotmat, workspace, g_verts, g_joints, need_pose=True, need_betas=True):
        pose = _f32(pose, 'pose')
        B = pose.shape[0]
        gv = _f32(g_verts, 'g_verts') if g_verts is not None else None
        gj = _f32(g_joints, 'g_joints') if g_joints is not None else None
        g_pose = torch.empty(B, 216 if is_rotmat else 72, device=pose.device, dtype=torch.float32) if need_pose else None
        g_betas = torch.empty(B, self.L, device=pose.device, dtype=torch.float32) if need_betas else None
        with torch.cuda.device(pose.device):
            check(lib().tuch_smpl_backward(self._h, _ptr(pose), int(bool(is_rotmat)), B, _ptr(workspace), _ptr(gv),
                                           _ptr(gj), _ptr(g_pose), _ptr(g_betas), _stream()), 'tuch_smpl_backward')
        return g_pose, g_betas


# ------------------------------------------------------------------ SMPLify-DC objective terms
PULL_THRESHOLD, PULL_ALL = 0, 1
REDUCE_SUM, REDUCE_MEAN = 0, 1


def reprojection_loss(joints, cam_t, center, joints_2d, conf, focal_length, sigma=100.0, cam_t_est=None,
                      depth_loss_weight=0.0, g_loss=None, want_grad=True):
    """losses.py:56-61 (+ the depth term of :146 when cam_t_est is given) -> dict(loss[B,J],
    depth[B] | None, g_joints[B,J,3] | None, g_cam_t[B,3] | None)."""
    joints, cam_t, center = _f32(joints, 'joints'), _f32(cam_t, 'cam_t'), _f32(center, 'center')
    joints_2d, conf = _f32(joints_2d, 'joints_2d'), _f32(conf, 'conf')
    B, J = joints.shape[0], joints.shape[1]
    if (tuple(joints.shape) != (B, J, 3) or tuple(cam_t.shape) != (B, 3) or tuple(center.shape) != (B, 2)
            or tuple(joints_2d.shape) != (B, J, 2) or tuple(conf.shape) != (B, J)):
        raise TuchError('reprojection_loss: inconsistent shapes %s %s %s %s %s' % (
            tuple(joints.shape), tuple(cam_t.shape), tuple(center.shape), tuple(joints_2d.shape), tuple(conf.shape)))
    est = _f32(cam_t_est, 'cam_t_est') if cam_t_est is not None else None
    gl = _f32(g_loss, 'g_loss') if g_loss is not None else None
    loss = torch.empty(B, J, device=joints.device, dtype=torch.float32)
    depth = torch.empty(B, device=joints.device, dtype=torch.float32) if est is not None else None
    gj = torch.empty(B, J, 3, device=joints.device, dtype=torch.float32) if want_grad else None
    gc = torch.empty(B, 3, device=joints.device, dtype=torch.float32) if want_grad else None
    with torch.cuda.device(joints.device):
        check(lib().tuch_reprojection_loss(_ptr(joints), _ptr(cam_t), _ptr(center), _ptr(joints_2d), _ptr(conf), B, J,
                                           float(focal_length), float(sigma), _ptr(est), float(depth_loss_weight),
                                           _ptr(gl), _ptr(loss), _ptr(depth), _ptr(gj), _ptr(gc), _stream()),
              'tuch_reprojection_loss')
    return dict(loss=loss, depth=depth, g_joints=gj, g_cam_t=gc)


class PriorHandle:
    """Device-resident max-mixture pose prior (prior.py:80-96 buffers)."""

    def __init__(self, means, precisions, nll_weights, device):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise TuchError('PriorHandle needs a CUDA device: tuch_b200 has no CPU fallback')
        f32 = lambda a: np.ascontiguousarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a),
                                             dtype=np.float32)
        mu, pr, w = f32(means), f32(precisions), f32(nll_weights).reshape(-1)
        self.M, self.D = int(mu.shape[0]), int(mu.shape[1])
        if tuple(pr.shape) != (self.M, self.D, self.D) or len(w) != self.M:
            raise TuchError('prior arrays have inconsistent shapes %s %s %s' % (mu.shape, pr.shape, w.shape))
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().tuch_prior_create(self.M, self.D, _hp(mu), _hp(pr), _hp(w), C.byref(self._h)), 'tuch_prior_create')

    def __del__(self):
        try:
            h = getattr(self, '_h', None)
            if h is not None and h.value:
                lib().tuch_prior_destroy(h)
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h


def pose_terms(prior, pose, betas=None, pose_prior_weight=1.0, angle_prior_weight=0.0, shape_prior_weight=0.0,
               want_grad=True):
    """prior.py:117-132 + losses.py:155-162 + the shape regulariser -> dict(value[B], prior[B],
    component[B], g_pose[B,D] | None, g_betas[B,L] | None).  Weights are the UNSQUARED ones."""
    pose = _f32(pose, 'pose')
    B, D = pose.shape
    be = _f32(betas, 'betas') if betas is not None else None
    L = be.shape[1] if be is not None else 0
    value = torch.empty(B, device=pose.device, dtype=torch.float32)
    pv = torch.empty(B, device=pose.device, dtype=torch.float32)
    comp = torch.empty(B, device=pose.device, dtype=torch.int32)
    gp = torch.empty(B, D, device=pose.device, dtype=torch.float32) if want_grad else None
    gb = torch.empty(B, L, device=pose.device, dtype=torch.float32) if (want_grad and be is not None) else None
    with torch.cuda.device(pose.device):
        check(lib().tuch_pose_terms(prior.handle if prior is not None else None, _ptr(pose), _ptr(be), B, D, L,
                                    float(pose_prior_weight), float(angle_prior_weight), float(shape_prior_weight),
                                    _ptr(value), _ptr(pv), _ptr(comp), _ptr(gp), _ptr(gb), _stream()), 'tuch_pose_terms')
    return dict(value=value, prior=pv, component=comp, g_pose=gp, g_betas=gb)


def contact_loss(points, argmin, exterior, euclthres, pull_mode=PULL_THRESHOLD, reduce_mode=REDUCE_SUM,
                 body_active=None, counts=None, weight=1.0, g_loss=None, g_points=None, want_parts=False):
    """Push/pull terms (losses.py:96-105 | loss.py:299-315 | eft/loss.py:158-166) -> (loss[B], parts | None).
    When g_points[B,N,3] is given, weight * g_loss[b] * d loss[b]/d points is ACCUMULATED into it."""
    p = _f32(points, 'points')
    B, N = p.shape[0], p.shape[1]
    am = _dev(argmin, 'argmin')
    if am.dtype != torch.int32:
        am = am.to(torch.int32)
    am = am.contiguous()
    ex = _dev(exterior, 'exterior').to(torch.uint8).contiguous()
    if tuple(am.shape) != (B, N) or tuple(ex.shape) != (B, N):
        raise TuchError('contact_loss: argmin / exterior must be [%d,%d]' % (B, N))
    act = _dev(body_active, 'body_active').to(torch.uint8).contiguous() if body_active is not None else None
    cnt = _dev(counts, 'counts').to(torch.int32).contiguous() if counts is not None else None
    gl = _f32(g_loss, 'g_loss') if g_loss is not None else None
    if g_points is not None and (g_points.dtype != torch.float32 or not g_points.is_contiguous()
                                 or tuple(g_points.shape) != (B, N, 3)):
        raise TuchError('contact_loss: g_points must be a contiguous fp32 [%d,%d,3] tensor' % (B, N))
    loss = torch.empty(B, device=p.device, dtype=torch.float32)
    parts = torch.empty(B, 4, device=p.device, dtype=torch.float32) if want_parts else None
    with torch.cuda.device(p.device):
        check(lib().tuch_contact_loss(_ptr(p), _ptr(am), _ptr(ex), _ptr(act), _ptr(cnt), B, N, float(euclthres),
                                      int(pull_mode), int(reduce_mode), float(weight), _ptr(gl), _ptr(loss),
                                      _ptr(parts), _ptr(g_points), _stream()), 'tuch_contact_loss')
    return loss, parts


def region_sum(verts, min_sq, arg_i, arg_j, body_active=None, weight=1.0, g_loss=None, g_verts=None):
    """r2r[b] = sum of the reported region minima (losses.py:116-117); optionally accumulates the
    gradient through the attaining vertex pairs into g_verts."""
    v = _f32(verts, 'verts')
    B, V = v.shape[0], v.shape[1]
    P = min_sq.shape[1]
    act = _dev(body_active, 'body_active').to(torch.uint8).contiguous() if body_active is not None else None
    gl = _f32(g_loss, 'g_loss') if g_loss is not None else None
    r2r = torch.empty(B, device=v.device, dtype=torch.float32)
    with torch.cuda.device(v.device):
        check(lib().tuch_region_sum(_ptr(v), B, V, P, _ptr(min_sq), _ptr(arg_i), _ptr(arg_j), _ptr(act), float(weight),
                                    _ptr(gl), _ptr(r2r), _ptr(g_verts), _stream()), 'tuch_region_sum')
    return r2r


def adam_step(param, grad, exp_avg, exp_avg_sq, step_dev, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """In-place torch.optim.Adam step on a contiguous fp32 CUDA tensor; step_dev is an int32 CUDA
    scalar holding the steps taken so far (incremented by the call)."""
    for t, n in ((param, 'param'), (grad, 'grad'), (exp_avg, 'exp_avg'), (exp_avg_sq, 'exp_avg_sq')):
        _dev(t, n)
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != param.numel():
            raise TuchError('adam_step: %s must be a contiguous fp32 tensor of %d elements' % (n, param.numel()))
    if step_dev.dtype != torch.int32 or not step_dev.is_cuda:
        raise TuchError('adam_step: step_dev must be an int32 CUDA tensor')
    with torch.cuda.device(param.device):
        check(lib().tuch_adam_step(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(),
                                   _ptr(step_dev), float(lr), float(beta1), float(beta2), float(eps), _stream()),
              'tuch_adam_step')


class ContactFitState:
    """The tensors of one SMPLify-DC stage-2 optimisation as tuch_contact_fit_step reads and writes them: the two
    parameter tensors and their torch.optim.Adam state (updated in place), the fixed inputs, the outputs of the last
    iteration.  step() enqueues ONE whole iteration (SMPL forward, contact_fitting_loss, backward, Adam) on the
    current stream through a single C-ABI call -- no torch op takes part."""

    def __init__(self, smpl, topo, prior, body_pose, global_orient, betas, camera_t, camera_center, joints_2d,
                 joints_conf, body_active=None, pair_active=None, euclthres=0.0, focal_length=5000.0, sigma=100.0,
                 pose_prior_weight=1.0, contact_loss_weight=1000.0, use_segments=True, lr=1e-2, betas_adam=(0.9, 0.999),
                 eps=1e-8, keep_grads=False):
        dev = body_pose.device
        B = body_pose.shape[0]
        self.smpl, self.topo, self.prior, self.B = smpl, topo, prior, B
        for t, shape, name in ((body_pose, (B, 69), 'body_pose'), (global_orient, (B, 3), 'global_orient'),
                               (betas, (B, smpl.L), 'betas'), (camera_t, (B, 3), 'camera_t'),
                               (camera_center, (B, 2), 'camera_center'), (joints_2d, (B, smpl.NO, 2), 'joints_2d'),
                               (joints_conf, (B, smpl.NO), 'joints_conf')):
            _dev(t, name)
            if t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != shape:
                raise TuchError('ContactFitState: %s must be a contiguous fp32 CUDA tensor %s, got %s %s'
                                % (name, shape, t.dtype, tuple(t.shape)))
        z = lambda *s, dt=torch.float32: torch.zeros(*s, device=dev, dtype=dt)
        self.body_pose, self.global_orient = body_pose, global_orient
        self.exp_avg_pose, self.exp_avg_sq_pose = z(B, 69), z(B, 69)
        self.exp_avg_orient, self.exp_avg_sq_orient = z(B, 3), z(B, 3)
        self.step_pose, self.step_orient = z((), dt=torch.int32), z((), dt=torch.int32)
        self.betas, self.camera_t, self.camera_center = betas, camera_t, camera_center
        self.joints_2d, self.joints_conf = joints_2d, joints_conf
        self.body_active = None if body_active is None else _dev(body_active, 'body_active').to(torch.uint8).contiguous()
        self.pair_active = None if pair_active is None else _dev(pair_active, 'pair_active').to(torch.uint8).contiguous()
        self.workspace = smpl.workspace(B)
        self.vertices, self.joints = z(B, smpl.V, 3), z(B, smpl.NO, 3)
        self.loss, self.per_body = z(()), z(B)
        self.exterior, self.argmin = z(B, smpl.V, dt=torch.uint8), z(B, smpl.V, dt=torch.int32)
        self.grad_body_pose = z(B, 69) if keep_grads else None
        self.grad_global_orient = z(B, 3) if keep_grads else None
        a = ContactFitArgs()
        for n in ('body_pose', 'global_orient', 'exp_avg_pose', 'exp_avg_sq_pose', 'exp_avg_orient', 'exp_avg_sq_orient',
                  'step_pose', 'step_orient', 'betas', 'camera_t', 'camera_center', 'joints_2d', 'joints_conf',
                  'body_active', 'pair_active', 'vertices', 'joints', 'loss', 'per_body', 'exterior', 'argmin',
                  'grad_body_pose', 'grad_global_orient'):
            t = getattr(self, n)
            setattr(a, n, t.data_ptr() if t is not None else None)
        a.smpl_workspace = self.workspace.data_ptr()
        a.euclthres, a.focal_length, a.sigma = float(euclthres), float(focal_length), float(sigma)
        a.pose_prior_weight, a.contact_loss_weight = float(pose_prior_weight), float(contact_loss_weight)
        a.use_segments = int(bool(use_segments))
        a.lr, a.beta1, a.beta2, a.eps = float(lr), float(betas_adam[0]), float(betas_adam[1]), float(eps)
        self._args = a

    def step(self):
        with torch.cuda.device(self.body_pose.device):
            check(lib().tuch_contact_fit_step(self.smpl.handle, self.topo.handle,
                                              self.prior.handle if self.prior is not None else None, self.B,
                                              C.byref(self._args), _stream()), 'tuch_contact_fit_step')
        return self.loss

    def reset_optimizer(self):
        for t in (self.exp_avg_pose, self.exp_avg_sq_pose, self.exp_avg_orient, self.exp_avg_sq_orient,
                  self.step_pose, self.step_orient):
            t.zero_()


# ------------------------------------------------------------------ f2 / f3: camera + pose bookkeeping
def estimate_translation(S, joints_2d, has_2d_kp_anno, focal_length=5000.0, img_size=224.0, n_openpose=25):
    """tuch/utils/geometry.py:156-205 on the device: S[B,J,3], joints_2d[B,J,3], has_2d_kp_anno[B] -> [B,3]."""
    S, k = _f32(S, 'S'), _f32(joints_2d, 'joints_2d')
    B, J = S.shape[0], S.shape[1]
    if k.shape != (B, J, 3) or S.shape[2] != 3:
        raise TuchError('estimate_translation: S and joints_2d must be [B,J,3]')
    has = has_2d_kp_anno.to(device=S.device, dtype=torch.uint8).contiguous()
    out = torch.empty(B, 3, device=S.device, dtype=torch.float32)
    with torch.cuda.device(S.device):
        check(lib().tuch_estimate_translation(_ptr(S), _ptr(k), _ptr(has), B, J, int(n_openpose), float(focal_length),
                                              float(img_size), _ptr(out), _stream()), 'tuch_estimate_translation')
    return out


def rotmat_to_angle_axis(rotmat):
    """torchgeometry.rotation_matrix_to_angle_axis: [N,3,3] or [N,3,4] -> [N,3]."""
    r = _f32(rotmat, 'rotmat')
    if r.dim() != 3 or r.shape[1] != 3 or r.shape[2] not in (3, 4):
        raise TuchError('rotmat_to_angle_axis: input must be [N,3,3] or [N,3,4], got %s' % (tuple(r.shape),))
    out = torch.empty(r.shape[0], 3, device=r.device, dtype=torch.float32)
    with torch.cuda.device(r.device):
        check(lib().tuch_rotmat_to_angle_axis(_ptr(r), r.shape[0], r.shape[2], _ptr(out), _stream()),
              'tuch_rotmat_to_angle_axis')
    return out


def fits_pose_transform(pose, rot_deg=None, is_flipped=None, flip_perm=None, flip_first=False):
    """FitsDict.rotate_pose / flip_pose (fits_dict.py:89-119) fused: pose[B,D] -> [B,D]."""
    p = _f32(pose, 'pose')
    B, D = p.shape
    rot = None if rot_deg is None else rot_deg.to(device=p.device, dtype=torch.float32).contiguous()
    fl = None if is_flipped is None else is_flipped.to(device=p.device, dtype=torch.uint8).contiguous()
    perm = None if flip_perm is None else flip_perm.to(device=p.device, dtype=torch.int32).contiguous()
    out = torch.empty_like(p)
    with torch.cuda.device(p.device):
        check(lib().tuch_fits_pose_transform(_ptr(p), _ptr(rot), _ptr(fl), _ptr(perm), B, D, int(bool(flip_first)),
                                             _ptr(out), _stream()), 'tuch_fits_pose_transform')
    return out
