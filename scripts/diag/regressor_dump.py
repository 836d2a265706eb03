"""GPU-side dump for the fp64 adjudication of RegressorLoss.contact_loss at SMPL size (tests/golden/
regressor_full_size.npz): per-HD-point selection, nearest point and inside flag of the product path in both
winding modes, plus loss and vertex gradient -> gpurun_out/<tag>_regdiag.npz (analysed on the CPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from tuch_b200 import synthetic as syn          # noqa: E402
import test_regressor_gpu as trg                # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else 'r2'
r = np.load(os.path.join(ROOT, 'tests', 'golden', 'regressor_full_size.npz'))
model = syn.make_lattice_body_model(seed=0)
geo = syn.make_geodesics(model['v_template'], model['faces'], cache_dir=os.environ.get('TUCH_B200_CACHE', '/tmp/tuch_b200_cache'))
a = dict(model=model, geo=geo, segs=syn.make_segments(model))
a['hd_reg'], a['hd_fidx'] = syn.make_hd_regressor(model, n_hd=int(r['n_hd']))
crit = trg.make_criterion(a, True, geothres=float(r['geothres']), B=2)
topo = crit._topo
out = {}
valid = torch.tensor([True, True], device='cuda:0')
for name, mode in (('fast', topo.WINDING_FAST), ('exact', topo.WINDING_EXACT)):
    topo.set_winding_mode(mode)
    pv = torch.tensor(r['verts'], device='cuda:0', requires_grad=True)
    val = crit.contact_loss(pv, valid)
    val.backward()
    g = torch.zeros_like(pv)
    loss, dbg = topo.regressor_contact_loss(pv.detach(), valid=valid, euclthres=0.02, use_hd=True, g_verts=g, debug=True)
    q = topo.contact_query(pv.detach(), use_segments=True)
    out.update({name + '/loss': val.item(), name + '/g_verts': pv.grad.cpu().numpy(), name + '/per_body': loss.cpu().numpy(),
                name + '/counts': dbg['counts'].cpu().numpy(), name + '/sel': dbg['sel'].cpu().numpy(),
                name + '/hd_argmin': dbg['hd_argmin'].cpu().numpy(), name + '/hd_exterior': dbg['hd_exterior'].cpu().numpy(),
                name + '/v_argmin': q['argmin'].cpu().numpy(), name + '/v_min_sq': q['min_sq'].cpu().numpy(),
                name + '/v_exterior': q['exterior'].cpu().numpy(), name + '/v_winding': q['winding'].cpu().numpy()})
    print(name, 'loss', val.item(), 'ref', float(r['loss']), 'rel', abs(val.item() - float(r['loss'])) / float(r['loss']))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, 'gpurun_out', tag + '_regdiag.npz'), **out)
