"""TEST INFRASTRUCTURE -- restatement of tuch/utils/segmentation.py:29-124 on numpy arrays.

A Segment owns: vidx (member vertex ids, segmentation.py:42), one vertex loop per boundary band
(:45-46), and the closed-segment face list = faces fully inside the segment + one cap fan per
band whose apex is an appended centroid vertex with index V + band_index (:51-66)."""
import numpy as np

from . import clib


class Segment:
    def __init__(self, name, vidx, bands, faces):
        self.name = name
        self.vidx = np.unique(np.asarray(vidx, dtype=np.int64))     # np.where(...) order, :42
        self.bands = [np.asarray(b, dtype=np.int64) for b in bands]
        faces = np.asarray(faces, dtype=np.int64)
        apex0 = int(faces.max()) + 1                                      # :36, :61
        inside = np.isin(faces, self.vidx).sum(1) == 3                    # :50
        caps = []
        for k, loop in enumerate(self.bands):
            for i in range(len(loop) - 1):                                # :62-63
                caps.append((loop[i + 1], loop[i], apex0 + k))
        self.faces = np.concatenate([faces[inside], np.asarray(caps, dtype=np.int64).reshape(-1, 3)], 0)

    def closed_triangles(self, v):
        """v[V,3] -> [F_seg,3,3]; appended apex = mean of the loop vertices as listed (:73-77)."""
        ext = [v] + [v[loop].mean(0, keepdims=True).astype(v.dtype) for loop in self.bands]
        return np.concatenate(ext, 0)[self.faces]

    def exterior(self, v, dtype=np.float32):
        """has_self_isect (:81-99): winding of the member vertices w.r.t. the closed segment <= 0.99."""
        v = np.asarray(v, dtype=dtype)
        return clib.winding_numbers(v[self.vidx], self.closed_triangles(v), dtype=dtype) <= 0.99


def build_segments(segdict, faces):
    """segdict: tuch_b200.synthetic.make_segments() output; order = dict order (segmentation.py:113-115)."""
    return [Segment(n, s['vidx'], list(s['bands'].values()), faces) for n, s in segdict.items()]
