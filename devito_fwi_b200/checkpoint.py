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
S ~ sqrt(2 * steps) minimises (2 * n_segments + S) slices of HBM. Spare HBM can be spent on keeping u.dt2 of
further trailing segments from pass 1 (``keep_segments=n`` or ``'auto'``), which are then not recomputed.

``forward(save='checkpoint')`` runs pass 1 while it records the receivers, and ``gradient(rec, u=<its result>)``
then only needs pass 2: forward + recompute + adjoint = 3 sweeps per shot gradient instead of the 4 of the
reference's call sequence (forward for the data, then forward + reverse inside the Revolver).
"""
import ctypes
import math

from . import _lib
from .sparse import sparse_map

__all__ = ['checkpointed_gradient', 'checkpointed_forward', 'CheckpointedWavefield', 'plan_segments']


def plan_segments(time_m, time_M, segment=None):
    steps = time_M - time_m + 1
    if steps <= 0:
        return []
    S = int(segment) if segment else max(1, int(math.ceil(math.sqrt(2.0 * steps))))
    return [(ta, min(ta + S - 1, time_M)) for ta in range(time_m, time_M + 1, S)]


class CheckpointedWavefield(object):
    """What ``AcousticWaveSolver.forward(save='checkpoint')`` returns in place of a saved TimeFunction:
    the checkpoints of pass 1 (two slices per segment) and u.dt2 of the last segment, so that
    ``gradient(rec, u=<this>)`` goes straight to pass 2 - the forward sweep that produced the synthetic
    data is not repeated (the reference's pyrevolve branch runs it a second time, wavesolver.py:188-201)."""
    save = None

    def __init__(self, solver, src, vp_dev, coef, dt, nt, time_m, time_M, segs, ring, ckpt, segbuf, nkeep, S):
        self.solver, self.src, self.vp_dev, self.coef, self.dt = solver, src, vp_dev, coef, dt
        self.nt, self.time_m, self.time_M, self.segs = nt, time_m, time_M, segs
        self.ring, self.ckpt, self.segbuf = ring, ckpt, segbuf
        self.nkeep, self.S = nkeep, S      # u.dt2 of the last nkeep segments is already in segbuf (S slices each)

    @property
    def nbytes(self):
        return sum(t.numel() * 4 for t in (self.ring, self.ckpt, self.segbuf))


HBM_FRACTION = 0.75     # share of the free HBM that stored u.dt2 segments may take (the rest: v, grad, records)


def _keep_segments(nseg, S, slice_bytes, reserve_slices=10):
    """How many trailing segments can keep their u.dt2 from pass 1 (each one saves a recompute sweep of S steps
    at the price of 4 B/pt more traffic in pass 1): as many as fit in HBM_FRACTION of what is free."""
    import torch
    free = torch.cuda.mem_get_info()[0] + torch.cuda.memory_reserved() - torch.cuda.memory_allocated()
    budget = HBM_FRACTION * free - reserve_slices * slice_bytes
    return int(max(1, min(nseg, budget // (S * slice_bytes))))


def checkpointed_forward(solver, src, rec, vp, dt, illum=None, **kwargs):
    """Pass 1: forward sweep on a 3-slot ring with receiver recording (and the source illumination),
    checkpointing the two live slices before every segment; the last ``keep_segments`` segments (default 1;
    ``'auto'``: as many as fit in HBM_FRACTION of the free HBM) store u.dt2 right away and are not recomputed by
    pass 2. Returns a CheckpointedWavefield."""
    import torch
    from .wavesolver import _ptr, _stream
    lib = _lib.lib()
    nt = min(rec.nt, src.nt) if rec is not None else src.nt
    time_m, time_M = solver._time_bounds(kwargs, nt)
    segs = plan_segments(time_m, time_M, kwargs.pop('segment', None))
    keep = kwargs.pop('keep_segments', None)
    grid = solver.model.grid
    g = solver._gs()
    vp_dev = solver._vp_dev(vp)
    coef = solver._coeffs(vp_dev, dt)
    src_map = sparse_map(grid, src.coordinates.data)
    src_dev = src._sdata.dev()
    rec_map = sparse_map(grid, rec.coordinates.data) if rec is not None else None
    rec_dev = rec._sdata.dev(write=True) if rec is not None else None
    illum_dev = illum._buf.dev(write=True) if illum is not None else None
    slice_shape = grid.slice_shape
    ring = torch.zeros((3,) + slice_shape, dtype=torch.float32, device='cuda')
    nseg = len(segs)
    S = max((tb - ta + 1) for ta, tb in segs) if segs else 1
    ckpt = torch.empty((max(nseg, 1), 2) + slice_shape, dtype=torch.float32, device='cuda')
    if keep is None:
        nkeep = 1                          # minimal footprint: only the last segment's u.dt2 comes from pass 1
    elif keep == 'auto':
        nkeep = _keep_segments(nseg, S, grid.slice_elems * 4)
    else:
        nkeep = max(1, min(int(keep), max(nseg, 1)))
    try:
        segbuf = torch.empty((nkeep * S,) + slice_shape, dtype=torch.float32, device='cuda')
    except torch.cuda.OutOfMemoryError:
        nkeep = 1
        segbuf = torch.empty((S,) + slice_shape, dtype=torch.float32, device='cuda')
    k0 = nseg - nkeep                      # first segment whose u.dt2 is kept
    cdt = ctypes.c_float(dt)
    for k, (ta, tb) in enumerate(segs):
        if k < k0:                          # kept segments are never restored: no checkpoint needed
            ckpt[k, 0].copy_(ring[(ta - 1) % 3])
            ckpt[k, 1].copy_(ring[ta % 3])
        _lib.check(lib.b2fwi_forward(
            ctypes.byref(g), _ptr(vp_dev), _ptr(coef), cdt, nt, ta, tb,
            _ptr(src_dev), src_map.byref(), _ptr(rec_dev), rec_map.byref() if rec_map is not None else None,
            _ptr(ring), 0, _ptr(illum_dev), _ptr(segbuf[(k - k0) * S:]) if k >= k0 else None, ta, _stream()))
    return CheckpointedWavefield(solver, src, vp_dev, coef, dt, nt, time_m, time_M, segs, ring, ckpt, segbuf,
                                 nkeep, S)


def checkpointed_gradient(solver, rec, v, grad, vp, dt, checkpoints=None, **kwargs):
    """Gradient with checkpointing; same results as ``jacobian_adjoint(rec, u_saved)``.
    ``checkpoints``: the CheckpointedWavefield of an earlier ``forward(save='checkpoint')`` with the same
    model; without it pass 1 is run here (the reference's behaviour)."""
    from .wavesolver import _ptr, _stream, _Timer, BYTES_ADJ, BYTES_FWD
    lib = _lib.lib()
    timer = _Timer(solver._profile)
    cw = checkpoints
    if cw is None:
        src = kwargs.pop('src', None) or solver.geometry.src
        nt_ = min(rec.nt, src.nt)
        tm, tM = solver._time_bounds(kwargs, nt_)
        cw = checkpointed_forward(solver, src, None, vp, dt, time_m=tm, time_M=tM, segment=kwargs.pop('segment', None),
                                  keep_segments=kwargs.pop('keep_segments', None))
        cw.nt = nt_
    nt, segs, ring, ckpt, segbuf = cw.nt, cw.segs, cw.ring, cw.ckpt, cw.segbuf
    grid = solver.model.grid
    g = solver._gs()
    vp_dev, coef, src = cw.vp_dev, cw.coef, cw.src
    src_map = sparse_map(grid, src.coordinates.data)
    rec_map = sparse_map(grid, rec.coordinates.data)
    src_dev = src._sdata.dev()
    rec_dev = rec._sdata.dev()
    v_dev = v._buf.dev(write=True)
    grad_dev = grad._buf.dev(write=True)
    nseg = len(segs)
    k0 = nseg - cw.nkeep
    cdt = ctypes.c_float(cw.dt)
    for k in range(nseg - 1, -1, -1):
        ta, tb = segs[k]
        if k >= k0:                         # u.dt2 kept from pass 1
            hist = segbuf[(k - k0) * cw.S:]
        else:                               # restore the checkpoint and recompute the segment (kept ones are consumed)
            hist = segbuf
            ring[(ta - 1) % 3].copy_(ckpt[k, 0])
            ring[ta % 3].copy_(ckpt[k, 1])
            _lib.check(lib.b2fwi_forward(
                ctypes.byref(g), _ptr(vp_dev), _ptr(coef), cdt, nt, ta, tb,
                _ptr(src_dev), src_map.byref(), None, None, _ptr(ring), 0, None, _ptr(hist), ta, _stream()))
        _lib.check(lib.b2fwi_gradient(
            ctypes.byref(g), _ptr(vp_dev), _ptr(coef), cdt, nt, ta, tb,
            _ptr(rec_dev), rec_map.byref(), _ptr(hist), 2, ta, _ptr(v_dev), _ptr(grad_dev), _stream()))
    steps = max(cw.time_M - cw.time_m + 1, 0)
    bpp = BYTES_ADJ + (1 if checkpoints is not None else 2) * BYTES_FWD
    summary = solver._summary('Gradient', timer.stop(), steps, bpp)
    return grad, summary
