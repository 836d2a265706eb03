"""Host-side mirror of tuch/utils/segmentation.py (BodySegment :29-99, BatchBodySegment :102-124).

Same constructor arguments, attributes (.names, .segmentation[name].segment_vidx, .segment_faces,
.bands_verts) and methods (has_self_isect, batch_has_self_isec, get_closed_segment); the closed
segment's winding test runs in the sm_100a winding kernel for the whole batch and all segments at
once instead of one PyTorch op chain + one device->host copy per segment per body.

Where the reference reads `smpl_segment_{name}.ply` through trimesh and the band loops from
`data.essentials.segments.smpl.segm_utils`, this mirror does the same when those are importable,
and otherwise accepts the same information as plain arrays via `segment_data`.
"""
import os.path as osp

import numpy as np
import torch
import torch.nn as nn

from .. import ops


def _load_segment_assets(name):
    """(segment_vidx, {band: vertex loop}) exactly as segmentation.py:40-46 obtains them."""
    try:
        import trimesh
        from configs import config
        from data.essentials.segments.smpl import segm_utils as exn
    except Exception as e:       # pragma: no cover - depends on the user's data tree
        raise ops.TuchError(
            'BodySegment(%r): the reference assets (trimesh, configs.config.SEGMENT_DIR, '
            'data.essentials.segments.smpl.segm_utils) are not importable (%s); pass segment_data=' % (name, e))
    path = osp.join(config.SEGMENT_DIR, 'smpl_segment_{}.ply'.format(name))
    mesh = trimesh.load(path, process=False)
    vidx = np.where(np.array(mesh.visual.vertex_colors[:, 0]) == 255)[0]
    return vidx, dict(exn.segments[name])


class BodySegment(nn.Module):
    def __init__(self, name, faces, append_idx=None, segment_data=None):
        super().__init__()
        self.device = faces.device
        self.name = name
        self.append_idx = faces.max().item() if append_idx is None else append_idx
        if segment_data is None:
            vidx, bands = _load_segment_assets(name)
        else:
            vidx, bands = segment_data['vidx'], segment_data['bands']
        self.segment_vidx = np.unique(np.asarray(vidx, dtype=np.int64))
        self.bands = list(bands.keys())
        self.bands_verts = [list(v) for v in bands.values()]
        self.bands_faces = self.create_band_faces().to(self.device)
        f = faces.squeeze().reshape(-1, 3)
        fn = f.detach().cpu().numpy()
        inside = np.where(np.isin(fn, self.segment_vidx).sum(1) == 3)[0]
        self.register_buffer('segment_faces', torch.cat((f[inside, :], self.bands_faces), 0))

    def create_band_faces(self):
        """cap fan per band: [loop[i+1], loop[i], apex], apex index = append_idx + 1 + band (:56-66)"""
        out = []
        for k, loop in enumerate(self.bands_verts):
            apex = self.append_idx + 1 + k
            out += [[loop[i + 1], loop[i], apex] for i in range(len(loop) - 1)]
        return torch.tensor(np.array(out, dtype=np.int64).reshape(-1, 3), dtype=torch.long)

    def topology_entry(self):
        return (self.name, self.segment_vidx, self.segment_faces.detach().cpu().numpy(), self.bands_verts)

    def _topology(self, num_verts):
        topo = getattr(self, '_topo', None)
        if topo is None or topo.V != num_verts or topo.device != self.segment_faces.device:
            topo = ops.Topology(np.zeros((0, 3), np.int64), num_verts, self.segment_faces.device)
            topo.set_segments([self.topology_entry()])
            self._topo = topo
        return topo

    def get_closed_segment(self, vertices):
        """[B, F_seg, 3, 3] closed-segment triangles (band centroids appended, :68-79)."""
        v = vertices.detach()
        ext = [v] + [v[:, loop, :].mean(1, keepdim=True) for loop in self.bands_verts]
        return torch.cat(ext, 1)[:, self.segment_faces]

    def has_self_isect(self, vertices):
        """exterior flags of the member vertices w.r.t. the closed segment (:81-99)."""
        flags, _ = self._topology(vertices.shape[1]).segment_exterior(vertices)
        return flags[0].squeeze()


class BatchBodySegment(nn.Module):
    def __init__(self, names, faces, segment_data=None):
        super().__init__()
        self.names = list(names)
        self.nv = faces.max().item()
        self.segmentation = {}
        for name in self.names:
            self.segmentation[name] = BodySegment(
                name, faces, segment_data=None if segment_data is None else segment_data[name])
        self._topo = None

    def topology_entries(self):
        return [self.segmentation[n].topology_entry() for n in self.names]

    def _topology(self, num_verts, device):
        if self._topo is None or self._topo.V != num_verts or self._topo.device != device:
            topo = ops.Topology(np.zeros((0, 3), np.int64), num_verts, device)
            topo.set_segments(self.topology_entries())
            self._topo = topo
        return self._topo

    def batch_has_self_isec(self, vertices):
        """list (self.names order) of exterior flags; one fused device pass (:117-124)."""
        flags, _ = self._topology(vertices.shape[1], vertices.device).segment_exterior(vertices)
        return [f.squeeze() for f in flags]
