"""On-device checkpoint + recompute driver for the gradient when the forward history does not
fit in HBM (3-D models).  Replaces pyrevolve / examples.checkpointing of the reference
(seismic/acoustic/wavesolver.py:188-201) with a two-level scheme that lives entirely in HBM:

  pass 1  forward sweep on a 3-slot ring; before each segment of S steps the two live slices are
          copied to a checkpoint (device-to-device). The LAST K steps run in history mode instead: the
          sweep writes u[t+1] into a K+2-slice buffer rather than the ring, which costs no extra traffic,
          so K is simply as large as the spare HBM allows (``keep='auto'``);
  pass 2  the kept steps first, then the segments in reverse order: restore the checkpoint into the first
          two slices of an S+2-slice buffer, recompute the segment in history mode (a PLAIN forward sweep:
          20 B per point, nothing but the wavefield is written), then run the adjoint + imaging sweep over
          the segment reading ONE history value per point.

The imaging sum is taken by parts (B2FWI_HIST_UVDT2, include/b2fwi.h):
    sum_t u.dt2[t] v[t]  =  sum_t u[t] v.dt2[t]  +  (u[M+1] v[M] - u[M] v[M+1] - u[m] v[m-1] + u[m-1] v[m]) / dt^2
with v.dt2 formed inside the adjoint sweep from the three adjoint levels it holds anyway. Storing u.dt2 instead
(round 1) made the recompute sweep write a second array: 76 B per point-step for the whole shot gradient against
~66 now. The boundary term vanishes for the FWI gradient (u[m-1] = u[m] = 0 and v starts from rest); for a
caller-supplied adjoint state ``v`` the M-part is added explicitly below.

The forward kernels are deterministic, so every segmentation gives the bitwise same gradient; against the
full-history imaging (u.dt2 from three saved slices) it differs by fp32 rounding only (tests/test_gpu_parity.py).

``forward(save='checkpoint')`` runs pass 1 while it records the receivers, and ``gradient(rec, u=<its result>)``
then only needs pass 2: forward + recompute + adjoint = 3 sweeps per shot gradient instead of the 4 of the
reference's call sequence (forward for the data, then forward + reverse inside the Revolver).
"""
import ctypes
import math

from . import _lib
from .sparse import sparse_map

__all__ = ['checkpointed_gradient', 'checkpointed_forward', 'CheckpointedWavefield', 'plan_segments', 'plan_keep']

HIST_UVDT2 = 3          # include/b2fwi.h


def plan_segments(time_m, time_M, segment=None):
    steps = time_M - time_m + 1
    if steps <= 0:
        return []
    S = int(segment) if segment else max(1, int(math.ceil(math.sqrt(2.0 * steps))))
    return [(ta, min(ta + S - 1, time_M)) for ta in range(time_m, time_M + 1, S)]


def plan_keep(steps, free_slices, segment=None, reserve_slices=8):
    """(K, S): how many trailing steps keep their wavefield from pass 1 and the segment length of the rest, for
    ``free_slices`` slices of spare HBM. Footprint: K + 2 (kept) + S + 2 (one recomputed segment) +
    2 * ceil((steps - K) / S) (checkpoints) + reserve (adjoint ring, gradient, records). Pure function (CPU tests)."""
    best = (0, int(segment) if segment else max(1, int(math.ceil(math.sqrt(2.0 * max(steps, 1))))))
    if steps <= 0:
        return best
    cands = [int(segment)] if segment else sorted(set(
        max(1, int(round(f * math.sqrt(2.0 * steps)))) for f in (0.5, 0.7, 0.85, 1.0, 1.2, 1.5, 2.0)))
    best_k = -1
    for S in cands:
        # largest K with K + 2 + S + 2 + 2*ceil((steps-K)/S) + reserve <= free_slices
        lo, hi = 0, steps
        while lo < hi:
            K = (lo + hi + 1) // 2
            rest = steps - K
            need = K + 2 + (S + 2 if rest > 0 else 0) + 2 * int(math.ceil(rest / float(S))) + reserve_slices
            if need <= free_slices:
                lo = K
            else:
                hi = K - 1
        if lo > best_k:
            best_k, best = lo, (lo, S)
    return best


class CheckpointedWavefield(object):
    """What ``AcousticWaveSolver.forward(save='checkpoint')`` returns in place of a saved TimeFunction:
    the checkpoints of pass 1 (two slices per segment) and the wavefield of the last K steps, so that
    ``gradient(rec, u=<this>)`` goes straight to pass 2 - the forward sweep that produced the synthetic
    data is not repeated (the reference's pyrevolve branch runs it a second time, wavesolver.py:188-201)."""
    save = None

    def __init__(self, solver, src, vp_dev, coef, dt, nt, time_m, time_M, segs, ring, ckpt, segbuf, keepbuf, tk, S):
        self.solver, self.src, self.vp_dev, self.coef, self.dt = solver, src, vp_dev, coef, dt
        self.nt, self.time_m, self.time_M, self.segs = nt, time_m, time_M, segs
        self.ring, self.ckpt, self.segbuf, self.keepbuf = ring, ckpt, segbuf, keepbuf
        self.tk, self.S = tk, S            # steps tk .. time_M live in keepbuf: slice i holds u[tk - 1 + i]

    @property
    def nkeep_steps(self):
        return max(self.time_M - self.tk + 1, 0)

    @property
    def nbytes(self):
        return sum(t.numel() * 4 for t in (self.ring, self.ckpt, self.segbuf, self.keepbuf) if t is not None)


import os

HBM_FRACTION = float(os.environ.get('B2FWI_HBM_FRACTION', 0.85))     # share of the free HBM the pass-1 history may take under keep='auto'


def _free_slices(slice_bytes):
    import torch
    free = torch.cuda.mem_get_info()[0] + torch.cuda.memory_reserved() - torch.cuda.memory_allocated()
    return int(HBM_FRACTION * free // slice_bytes)


def _hist_ptr(buf, t0, slice_elems):
    """Pointer such that (ptr + t * slice) is the slice of time level t when buf[0] holds level t0
    (b2fwi_forward save=1 addresses u + t * elems; only levels >= t0 are touched)."""
    return ctypes.c_void_p(buf.data_ptr() - int(t0) * int(slice_elems) * 4)


def checkpointed_forward(solver, src, rec, vp, dt, illum=None, **kwargs):
    """Pass 1: forward sweep with receiver recording (and the source illumination), checkpointing the two live
    slices before every segment; the last K steps write the wavefield into the history buffer of pass 2.
    ``keep_segments``: None / 'auto' = K as large as HBM_FRACTION of the free HBM allows, an int n = the last n
    segments, 0 = minimal footprint (the last segment only). Returns a CheckpointedWavefield."""
    import torch
    from .wavesolver import _ptr, _stream
    lib = _lib.lib()
    nt = min(rec.nt, src.nt) if rec is not None else src.nt
    time_m, time_M = solver._time_bounds(kwargs, nt)
    segment = kwargs.pop('segment', None)
    keep = kwargs.pop('keep_segments', None)
    steps = max(time_M - time_m + 1, 0)
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
    elems = grid.slice_elems
    ring = torch.zeros((3,) + slice_shape, dtype=torch.float32, device='cuda')
    if keep is None or keep == 'auto':
        K, S = plan_keep(steps, _free_slices(elems * 4), segment)
    else:
        S = int(segment) if segment else max(1, int(math.ceil(math.sqrt(2.0 * max(steps, 1)))))
        K = min(steps, max(1, int(keep)) * S)
    K = max(K, min(S, steps))                       # at least the last segment comes from pass 1
    while True:
        tk = time_M - K + 1                          # first kept step
        segs = plan_segments(time_m, tk - 1, S)
        try:
            ckpt = torch.empty((len(segs), 2) + slice_shape, dtype=torch.float32, device='cuda') if segs else None
            segbuf = torch.empty((S + 2,) + slice_shape, dtype=torch.float32, device='cuda') if segs else None
            keepbuf = torch.empty((K + 2,) + slice_shape, dtype=torch.float32, device='cuda') if K > 0 else None
            break
        except torch.cuda.OutOfMemoryError:
            ckpt = segbuf = keepbuf = None
            torch.cuda.empty_cache()
            if K <= min(S, steps):
                raise
            K = max(min(S, steps), K // 2)
    cdt = ctypes.c_float(dt)

    def sweep(ta, tb, u_ptr, save, ill):
        _lib.check(lib.b2fwi_forward(
            ctypes.byref(g), _ptr(vp_dev), _ptr(coef), cdt, nt, ta, tb,
            _ptr(src_dev), src_map.byref(), _ptr(rec_dev), rec_map.byref() if rec_map is not None else None,
            u_ptr, save, ill, None, 0, _stream()))

    for k, (ta, tb) in enumerate(segs):
        ckpt[k, 0].copy_(ring[(ta - 1) % 3])
        ckpt[k, 1].copy_(ring[ta % 3])
        sweep(ta, tb, _ptr(ring), 0, _ptr(illum_dev))
    if K > 0:
        keepbuf[0].copy_(ring[(tk - 1) % 3])
        keepbuf[1].copy_(ring[tk % 3])
        sweep(tk, time_M, _hist_ptr(keepbuf, tk - 1, elems), 1, _ptr(illum_dev))
    return CheckpointedWavefield(solver, src, vp_dev, coef, dt, nt, time_m, time_M, segs, ring, ckpt, segbuf,
                                 keepbuf, tk, S)


def checkpointed_gradient(solver, rec, v, grad, vp, dt, checkpoints=None, v_from_rest=True, **kwargs):
    """Gradient with checkpointing; same results as ``jacobian_adjoint(rec, u_saved)`` up to fp32 rounding.
    ``checkpoints``: the CheckpointedWavefield of an earlier ``forward(save='checkpoint')`` with the same
    model; without it pass 1 is run here (the reference's behaviour). ``v_from_rest=False``: the caller's ``v``
    holds a non-zero adjoint state, whose boundary term of the summation by parts is added here."""
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
    nt, segs, ckpt, segbuf, keepbuf = cw.nt, cw.segs, cw.ckpt, cw.segbuf, cw.keepbuf
    grid = solver.model.grid
    elems = grid.slice_elems
    g = solver._gs()
    vp_dev, coef, src = cw.vp_dev, cw.coef, cw.src
    src_map = sparse_map(grid, src.coordinates.data)
    rec_map = sparse_map(grid, rec.coordinates.data)
    src_dev = src._sdata.dev()
    rec_dev = rec._sdata.dev()
    v_dev = v._buf.dev(write=True)
    grad_dev = grad._buf.dev(write=True)
    cdt = ctypes.c_float(cw.dt)

    # kernel='OT4': the sweep's own result is not final (the double-Laplacian term is added after it), so the imaging
    # reads u.dt2 off three slices of the segment's wavefield instead (B2FWI_HIST_U; the buffers hold ta-1 .. tb+1)
    ot4 = solver.kernel == 'OT4'
    kind = 1 if ot4 else HIST_UVDT2

    def adjoint(ta, tb, hist, t0):
        _lib.check(lib.b2fwi_gradient(
            ctypes.byref(g), _ptr(vp_dev), _ptr(coef), cdt, nt, ta, tb,
            _ptr(rec_dev), rec_map.byref(), _ptr(hist), kind, t0, _ptr(v_dev), _ptr(grad_dev), _stream()))

    if keepbuf is not None and cw.time_M >= cw.tk:
        if not v_from_rest and not ot4:
            # grad += -(u[M+1] v[M] - u[M] v[M+1]) / dt^2 ; the m-part is zero: pass 1 starts from a zero ring
            M = cw.time_M
            inv_dt2 = 1.0 / (float(cw.dt) * float(cw.dt))
            grad_dev.add_((keepbuf[-1] * v_dev[M % 3] - keepbuf[-2] * v_dev[(M + 1) % 3]) * (-inv_dt2))
        adjoint(cw.tk, cw.time_M, keepbuf, cw.tk - 1)
    for k in range(len(segs) - 1, -1, -1):
        ta, tb = segs[k]
        segbuf[0].copy_(ckpt[k, 0])
        segbuf[1].copy_(ckpt[k, 1])
        _lib.check(lib.b2fwi_forward(
            ctypes.byref(g), _ptr(vp_dev), _ptr(coef), cdt, nt, ta, tb,
            _ptr(src_dev), src_map.byref(), None, None, _hist_ptr(segbuf, ta - 1, elems), 1, None, None, 0,
            _stream()))
        adjoint(ta, tb, segbuf, ta - 1)
    steps = max(cw.time_M - cw.time_m + 1, 0)
    bpp = BYTES_ADJ + (1 if checkpoints is not None else 2) * BYTES_FWD
    summary = solver._summary('Gradient', timer.stop(), steps, bpp)
    return grad, summary
