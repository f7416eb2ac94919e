"""The named configurations of BASELINE.json, rebuilt from the reference's driver scripts
(circle_fwi.py:62-96, marmousi_fwi.py:62-117, marmousi2_fwi.py:61-103, acoustic_example.py:26-63)."""
import os

import numpy as np

from .model import Model
from .geometry import AcquisitionGeometry, setup_geometry
from .preset_models import demo_model

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')

__all__ = ['marmousi', 'marmousi2', 'circle', 'layered3d', 'load_vp']


def load_vp(dataset, name, shape):
    return np.fromfile(os.path.join(_DATA, dataset, name), dtype=np.float32).reshape(shape) / 1000


def _survey(models, shape, spacing, nsrc, t0, tn, f0):
    any_model = models[0]
    src = np.empty((nsrc, 2))
    src[:, 0] = np.linspace(0, any_model.domain_size[0], num=nsrc)
    src[:, -1] = 2 * spacing[0]
    rec = np.empty((shape[0], 2))
    rec[:, 0] = np.linspace(spacing[0], any_model.domain_size[0] - spacing[0], num=shape[0])
    rec[:, 1] = 2 * spacing[0]
    return [AcquisitionGeometry(m, rec, src, t0, tn, f0=f0, src_type='Ricker') for m in models]


def marmousi(nsrc=29, tn=4000.):
    """marmousi_fwi.py: returns geometries of (true, smooth_20 initial, constant 1.5) models + bathy mask."""
    shape, spacing, so, nbl, dt = (300, 106), (30., 30.), 8, 40, 2.95
    vps = [load_vp('SMARMN', 'vp.true', shape), load_vp('SMARMN', 'vp.smooth_20', shape),
           np.ones(shape) * 1.5]
    models = [Model(origin=(0, 0), spacing=spacing, shape=shape, space_order=so, vp=v, nbl=nbl,
                    fs=False, dt=dt) for v in vps]
    mask = np.ones(shape, dtype=np.float32)
    mask[:, :7] = 0
    return _survey(models, shape, spacing, nsrc, 0., tn, 0.007) + [mask]


def marmousi2(nsrc=31, tn=4500.):
    """marmousi2_fwi.py: same structure on SMARM2 (340x140, dt=3.0)."""
    shape, spacing, so, nbl, dt = (340, 140), (30., 30.), 8, 40, 3.0
    vps = [load_vp('SMARM2', 'vp.true', shape), load_vp('SMARM2', 'vp.smooth_20', shape),
           np.ones(shape) * 1.5]
    models = [Model(origin=(0, 0), spacing=spacing, shape=shape, space_order=so, vp=v, nbl=nbl,
                    fs=False, dt=dt) for v in vps]
    mask = np.ones(shape, dtype=np.float32)
    mask[:, :15] = 0
    return _survey(models, shape, spacing, nsrc, 0., tn, 0.007) + [mask]


def circle(space_order=6, nsrc=11):
    """circle_fwi.py: transmission set-up through a disc anomaly; returns (true, initial) geometries."""
    shape, spacing, nbl, dt, radius = (201, 201), (10., 10.), 40, 1., 60
    kw = dict(origin=(0, 0), shape=shape, spacing=spacing, space_order=space_order, nbl=nbl, dt=dt)
    true_model = demo_model('circle-isotropic', vp_circle=3.6, vp_background=3, r=radius, **kw)
    init_model = demo_model('circle-isotropic', vp_circle=3, vp_background=3, r=radius, **kw)
    src = np.empty((nsrc, 2))
    src[:, 1] = np.linspace(0, true_model.domain_size[0], num=nsrc)
    src[:, 0] = 20.
    rec = np.empty((shape[0], 2))
    rec[:, 1] = np.linspace(spacing[0], true_model.domain_size[0] - spacing[0], num=shape[0])
    rec[:, 0] = 1980.
    return [AcquisitionGeometry(m, rec, src, 0., 1000., f0=0.010, src_type='Ricker')
            for m in (true_model, init_model)]


def layered3d(n=512, space_order=8, nbl=40, tn=1250., spacing=15.0, rec_decimate=1):
    """3-D layered model of acoustic_example.py (layers-isotropic, 1.5/2.5/3.5 km/s)."""
    model = demo_model('layers-isotropic', space_order=space_order, shape=(n, n, n), nbl=nbl,
                       spacing=(spacing,) * 3)
    geometry = setup_geometry(model, tn)
    if rec_decimate > 1:
        keep = geometry.rec_positions.reshape(n, n, 3)[::rec_decimate, ::rec_decimate].reshape(-1, 3)
        geometry = AcquisitionGeometry(model, keep, geometry.src_positions, 0.0, tn, src_type='Ricker',
                                       f0=0.010)
    return geometry
