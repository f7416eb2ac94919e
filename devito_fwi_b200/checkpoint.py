"""On-device checkpoint + recompute driver for the gradient when the forward history does not
fit in HBM (3-D models).  Replaces pyrevolve / examples.checkpointing of the reference
(seismic/acoustic/wavesolver.py:188-201) with a two-level scheme that lives entirely in HBM:

  pass 1  forward sweep on a 3-slot ring; before each segment of S steps the two live slices are
          copied to a checkpoint (device-to-device);
  pass 2  segments in reverse order: restore the checkpoint, recompute the segment while the
          forward kernel also stores u.dt2 per step into an S-slice buffer, then run the
          adjoint + imaging sweep over the segment reading one history value per point.

The forward kernels are deterministic, so the recomputed wavefield - and therefore the gradient -
is bitwise identical to the one obtained from a full saved history (tests/test_gpu_parity.py).
S ~ sqrt(2 * steps) minimises (2 * n_segments + S) slices of HBM.
"""
import ctypes
import math

from . import _lib
from .sparse import sparse_map

__all__ = ['checkpointed_gradient', 'plan_segments']


def plan_segments(time_m, time_M, segment=None):
    steps = time_M - time_m + 1
    if steps <= 0:
        return []
    S = int(segment) if segment else max(1, int(math.ceil(math.sqrt(2.0 * steps))))
    return [(ta, min(ta + S - 1, time_M)) for ta in range(time_m, time_M + 1, S)]


def checkpointed_gradient(solver, rec, v, grad, vp, dt, **kwargs):
    """Gradient with checkpointing; same results as ``jacobian_adjoint(rec, u_saved)``."""
    import torch
    from .wavesolver import _ptr, _stream, _Timer, BYTES_ADJ, BYTES_FWD
    lib = _lib.lib()
    src = kwargs.pop('src', None) or solver.geometry.src
    nt = min(rec.nt, src.nt)
    time_m, time_M = solver._time_bounds(kwargs, nt)
    segs = plan_segments(time_m, time_M, kwargs.pop('segment', None))
    grid = solver.model.grid
    g = solver._gs()
    vp_dev = solver._vp_dev(vp)
    coef = solver._coeffs(vp_dev, dt)
    src_map = sparse_map(grid, src.coordinates.data)
    rec_map = sparse_map(grid, rec.coordinates.data)
    src_dev = src._sdata.dev()
    rec_dev = rec._sdata.dev()
    v_dev = v._buf.dev(write=True)
    grad_dev = grad._buf.dev(write=True)
    slice_shape = grid.slice_shape
    ring = torch.zeros((3,) + slice_shape, dtype=torch.float32, device='cuda')
    nseg = len(segs)
    S = max((tb - ta + 1) for ta, tb in segs) if segs else 1
    ckpt = torch.empty((max(nseg, 1), 2) + slice_shape, dtype=torch.float32, device='cuda')
    segbuf = torch.empty((S,) + slice_shape, dtype=torch.float32, device='cuda')
    cdt = ctypes.c_float(dt)
    timer = _Timer(solver._profile)

    def fwd(ta, tb, d2u):
        _lib.check(lib.b2fwi_forward(
            ctypes.byref(g), _ptr(vp_dev), _ptr(coef), cdt, nt, ta, tb,
            _ptr(src_dev), src_map.byref(), None, None, _ptr(ring), 0, None,
            _ptr(segbuf) if d2u else None, ta, _stream()))

    for k, (ta, tb) in enumerate(segs):
        ckpt[k, 0].copy_(ring[(ta - 1) % 3])
        ckpt[k, 1].copy_(ring[ta % 3])
        fwd(ta, tb, d2u=(k == nseg - 1))
    for k in range(nseg - 1, -1, -1):
        ta, tb = segs[k]
        if k != nseg - 1:
            ring[(ta - 1) % 3].copy_(ckpt[k, 0])
            ring[ta % 3].copy_(ckpt[k, 1])
            fwd(ta, tb, d2u=True)
        _lib.check(lib.b2fwi_gradient(
            ctypes.byref(g), _ptr(vp_dev), _ptr(coef), cdt, nt, ta, tb,
            _ptr(rec_dev), rec_map.byref(), _ptr(segbuf), 2, ta, _ptr(v_dev), _ptr(grad_dev), _stream()))
    steps = max(time_M - time_m + 1, 0)
    summary = solver._summary('Gradient', timer.stop(), steps, BYTES_ADJ + 2 * BYTES_FWD)
    return grad, summary
