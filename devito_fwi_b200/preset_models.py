"""Preset isotropic-acoustic models used by the drivers and KATs
(reference: seismic/preset_models.py:10-126,231-251)."""
import numpy as np

from .model import SeismicModel

__all__ = ['demo_model']


def demo_model(preset, **kwargs):
    """``constant-isotropic``, ``layers-isotropic`` (n horizontal layers, 1.5 .. 3.5 km/s) and
    ``circle-isotropic`` (disc anomaly in a constant background)."""
    space_order = kwargs.pop('space_order', 2)
    shape = kwargs.pop('shape', (101, 101))
    spacing = kwargs.pop('spacing', tuple([10. for _ in shape]))
    origin = kwargs.pop('origin', tuple([0. for _ in shape]))
    nbl = kwargs.pop('nbl', 10)
    dtype = kwargs.pop('dtype', np.float32)
    vp = kwargs.pop('vp', 1.5)
    nlayers = kwargs.pop('nlayers', 3)
    fs = kwargs.pop('fs', False)
    name = preset.lower()

    if name in ('constant-isotropic', 'constant'):
        v = vp
    elif name == 'layers-isotropic':
        vp_top = kwargs.pop('vp_top', 1.5)
        vp_bottom = kwargs.pop('vp_bottom', 3.5)
        v = np.full(shape, vp_top, dtype=dtype)
        layer_v = np.linspace(vp_top, vp_bottom, nlayers)
        thickness = int(shape[-1] / nlayers)
        for i in range(1, nlayers):
            v[..., i * thickness:] = layer_v[i]
    elif name == 'circle-isotropic':
        vp_circle = kwargs.pop('vp_circle', 3.0)
        vp_background = kwargs.pop('vp_background', 2.5)
        r = kwargs.pop('r', 15)
        assert len(shape) == 2
        v = np.full(shape, vp_background, dtype=dtype)
        a, b = shape[0] / 2, shape[1] / 2
        y, x = np.ogrid[-a:shape[0]-a, -b:shape[1]-b]
        v[x*x + y*y <= r*r] = vp_circle
    else:
        raise NotImplementedError("preset `%s` is outside the isotropic-acoustic hot path" % preset)

    return SeismicModel(space_order=space_order, vp=v, origin=origin, shape=shape, dtype=dtype,
                        spacing=spacing, nbl=nbl, bcs="damp", fs=fs, **kwargs)
