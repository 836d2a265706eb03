import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
CACHE = os.environ.get('TUCH_B200_CACHE', '/tmp/tuch_b200_cache')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope='session')
def small_assets():
    """The asset set tests/golden/make_golden.py used (V=122, F=240), regenerated deterministically."""
    from tuch_b200 import synthetic as syn
    d = golden('assets_digest.npz')
    model = syn.make_body_model(int(d['rings']), int(d['segs']), seed=0)
    geo = syn.make_geodesics(model['v_template'], model['faces'])
    regions = syn.make_regions(model, max_pairs=12)
    segs = syn.make_segments(model)
    hd_reg, hd_fidx = syn.make_hd_regressor(model, n_hd=300)
    gmm = syn.make_gmm()
    assert abs(model['v_template'].astype(np.float64).sum() - float(d['v_sum'])) < 1e-9
    assert abs(geo.astype(np.float64).sum() - float(d['geo_sum'])) < 1e-6 * abs(float(d['geo_sum']))
    assert len(regions['classes']) == int(d['n_classes'])
    return dict(model=model, geo=geo, regions=regions, segs=segs, hd_reg=hd_reg, hd_fidx=hd_fidx, gmm=gmm)


def _full(model):
    from tuch_b200 import synthetic as syn
    geo = syn.make_geodesics(model['v_template'], model['faces'], cache_dir=CACHE)
    return dict(model=model, geo=geo, regions=syn.make_regions(model), segs=syn.make_segments(model), gmm=syn.make_gmm())


@pytest.fixture(scope='session')
def full_assets():
    """SMPL-sized synthetic assets (V=6890, F=13776) with SMPL-like triangle statistics (lattice body);
    the geodesic matrix is cached under CACHE."""
    from tuch_b200 import synthetic as syn
    return _full(syn.make_lattice_body_model(seed=0))


@pytest.fixture(scope='session')
def full_assets_uv():
    """Same counts on the UV 'starfish' tessellation (long sliver triangles): the stress case for
    anything that clusters or bounds triangles spatially."""
    from tuch_b200 import synthetic as syn
    return _full(syn.make_body_model(84, 82, seed=0))
