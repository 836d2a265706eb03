"""TEST INFRASTRUCTURE -- restatement of tuch/train/loss.py:30-41 (batch_face_normals),
:240-317 (RegressorLoss.contact_loss incl. the HD-point path) and
tuch/train/train_module.py:69-91 (contact_from_verts)."""
import numpy as np
import torch

from . import clib


def face_normals(tris):
    """loss.py:30-41 on [F,3,3] numpy."""
    n = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0])
    return n / np.linalg.norm(n, axis=1, keepdims=True)


def regressor_contact_loss(pred_vertices, valid_fit, faces, geomask_np, euclthres, segments,
                           hd_regressor=None, hd_face_idx=None, use_hd=True, return_aux=False):
    """loss.py:240-317.  pred_vertices[B,V,3] torch (autograd); faces numpy [F,3];
    hd_regressor numpy [N,V]; hd_face_idx numpy [N] (faces_vert_is_sampled_from)."""
    B = pred_vertices.shape[0]
    dt = np.float64 if pred_vertices.dtype == torch.float64 else np.float32
    per_body = [pred_vertices.new_zeros(()) for _ in range(B)]
    aux = {}
    hd_first_vertex = None if hd_face_idx is None else faces[hd_face_idx][:, 0]     # loss.py:89
    for b in np.where(np.asarray(valid_fit))[0]:
        vb = pred_vertices[b]
        v = vb.detach().numpy().astype(dt)
        tris = v[faces]
        ext = clib.winding_numbers(v, tris, dtype=dt) <= 0.99                        # :260-262
        for seg in segments:                                                         # :264-266 (always)
            ext[seg.vidx[~seg.exterior(v, dtype=dt)]] = True
        am, mn = clib.masked_nearest(v, geomask_np, dtype=dt)                        # :269-270
        if use_hd:
            sel_v = np.where((mn < euclthres ** 2) | ~ext)[0]                        # :278
            sel_f = np.where(np.isin(faces, sel_v).any(1))[0]                        # :279-280
            sel_hd = np.isin(hd_face_idx, sel_f)                                     # :281
            aux[int(b)] = dict(sel_hd=sel_hd, exterior_v=ext.copy(), argmin_v=am)
            if sel_hd.sum() == 0:                                                    # :284, :300-301
                continue
            R = torch.as_tensor(hd_regressor[sel_hd], dtype=pred_vertices.dtype)
            hd = R @ vb                                                              # :285
            hdn = hd.detach().numpy().astype(dt)
            gv = hd_first_vertex[sel_hd]
            hd_geo = geomask_np[gv][:, gv]                                           # :289
            ham, _ = clib.masked_nearest(hdn, hd_geo, dtype=dt)                      # :288-291
            off = hdn + (0.001 * face_normals(tris)[hd_face_idx[sel_hd]]).astype(dt)  # :295-296
            hext = clib.winding_numbers(off, tris, dtype=dt) <= 0.99                 # :297
            d = torch.norm(hd - hd[torch.as_tensor(ham, dtype=torch.long)], dim=1)   # :299
            e = torch.as_tensor(hext)
            aux[int(b)].update(hd_exterior=hext, hd_argmin=ham)
        else:
            d = torch.norm(vb - vb[torch.as_tensor(am, dtype=torch.long)], dim=1)    # :303
            e = torch.as_tensor(ext)
        pull = (0.005 * torch.tanh(d[e] / 0.005) ** 2).sum() if e.any() else d.new_zeros(())     # :307-308
        push = (torch.tanh(d[~e] / 0.04) ** 2).sum() if (~e).any() else d.new_zeros(())          # :311-312
        per_body[int(b)] = pull + push
    per_body = torch.stack(per_body)
    out = per_body[torch.as_tensor(np.asarray(valid_fit))].mean()                                # :317
    return (out, aux) if return_aux else out


def contact_from_verts(verts, cdict):
    """train_module.py:69-91: [B, n_pairs] min squared (expansion-form) distance per region pair."""
    v = np.asarray(verts, np.float32)
    out = np.zeros((v.shape[0], len(cdict['classes'])), np.float32)
    for b in range(v.shape[0]):
        for k, (ra, rb) in enumerate(cdict['classes']):
            out[b, k] = clib.region_min(v[b], None, cdict['csig'][ra], cdict['csig'][rb])[0]
    return out
