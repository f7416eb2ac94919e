"""Helpers shared by the parity tests: build the oracle's numpy model from a product Model."""
import numpy as np

from oracle import ref


def ref_model(model, dtype=np.float64):
    nbl = model.nbl
    sl = tuple(slice(nbl, -nbl) if nbl else slice(None) for _ in model.shape)
    vp = np.array(model.vp.data)[sl] if hasattr(model.vp.data, 'shape') and np.ndim(model.vp.data) else \
        np.full(model.shape, float(model.vp.data))
    rm = ref.RefModel([float(o) for o in model.origin], [float(s) for s in model.spacing], model.shape,
                      model.space_order, vp.astype(dtype), nbl=nbl, dtype=dtype, dt=model._dt)
    # the product model may have been updated on the padded grid (e.g. gradient-descent step)
    if np.ndim(model.vp.data):
        rm.vp = np.array(model.vp.data, dtype=dtype)
    return rm


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)
