"""AcousticWaveSolver: the reference's propagator API on hand-written sm_100a kernels.

Signatures, defaults, ownership rules and return values follow
seismic/acoustic/wavesolver.py:28-246 of the reference: caller-supplied ``rec``/``u``/``v``/``grad``/``srca``
are mutated in place and returned, ``u``/``v`` supply the initial state, ``grad`` is accumulated into
(``Inc``, operators.py:217), ``vp`` may be a Function, a Constant or a float, ``dt=`` / ``time_m=`` /
``time_M=`` are honoured and other Operator.apply kwargs (``autotune``, ``opt``, ...) are ignored.
Every call goes through the C ABI of libb2fwi.so (include/b2fwi.h); there is no other compute path.
"""
import ctypes
import time as _time

import numpy as np

from . import _lib
from .grid import Function, TimeFunction, Constant, HALO
from .sparse import sparse_map

__all__ = ['AcousticWaveSolver', 'PerformanceSummary']

BYTES_FWD = 20   # algorithmic bytes per grid-point-step
BYTES_ADJ = 32


class PerformanceSummary(dict):
    """Stand-in for devito's PerformanceSummary: ``timings``, ``gpointss``, ``gflopss``, ``oi``
    plus achieved algorithmic GB/s of the sweep (CUDA-event timed)."""

    def __init__(self, name, seconds, points, steps, bytes_per_point, flops_per_point):
        super(PerformanceSummary, self).__init__()
        self.name = name
        self.time = seconds
        self.timings = {'section0': seconds}
        work = float(points) * steps
        self.gpointss = work / seconds / 1e9 if seconds > 0 else 0.0
        self.gflopss = self.gpointss * flops_per_point
        self.gbytess = self.gpointss * bytes_per_point
        self.oi = flops_per_point / float(bytes_per_point)
        self[name] = self.timings

    def __repr__(self):
        return "PerformanceSummary(%s: %.6f s, %.2f Gpts/s, %.1f GB/s)" % (
            self.name, self.time, self.gpointss, self.gbytess)


def grid_struct(grid, space_order, kernel='OT2'):
    g = _lib.Grid()
    g.kernel = 1 if kernel == 'OT4' else 0
    g.ndim = grid.dim
    g.space_order = int(space_order)
    g.halo = HALO
    g.fs = 1 if getattr(grid, 'fs', False) else 0
    for d in range(grid.dim):
        g.shape[d] = grid.shape[d]
        g.spacing[d] = float(grid.spacing[d])
        g.origin[d] = float(grid.origin[d])
    return g


def _stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class _Timer(object):
    def __init__(self, enabled=True):
        import torch
        self.enabled = enabled
        if enabled:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def stop(self):
        if not self.enabled:
            return 0.0
        self.e1.record()
        self.e1.synchronize()
        return self.e0.elapsed_time(self.e1) * 1e-3


class AcousticWaveSolver(object):
    """Forward / adjoint / gradient operators of the isotropic acoustic wave equation.

    Parameters as in the reference (wavesolver.py:28): ``model``, ``geometry``, ``kernel`` ('OT2', or
    'OT4' on the streaming engine), ``space_order`` (2..16, even). ``profile=False`` skips the CUDA-event
    timing and its stream synchronisation (summary is then ``None``).
    """

    def __init__(self, model, geometry, kernel='OT2', space_order=4, **kwargs):
        self.model = model
        self.model._initialize_bcs(bcs="damp")
        self.geometry = geometry

        assert self.model.grid == geometry.grid

        if kernel not in ('OT2', 'OT4'):
            raise ValueError("Unrecognized kernel")          # operators.py:52-53
        if kernel == 'OT4' and getattr(model, 'fs', False):
            raise NotImplementedError("kernel='OT4' with a free surface is not implemented")
        if space_order % 2 or not 2 <= space_order <= 16:
            raise ValueError("space_order must be even and in [2, 16]")
        if np.dtype(model.dtype) != np.float32:
            raise NotImplementedError("device compute is float32 only (the reference's default dtype)")
        self.space_order = space_order
        self.kernel = kernel
        self._profile = kwargs.pop('profile', True)
        self._kwargs = kwargs
        self._coef = None
        self._coef_key = None

    @property
    def dt(self):
        # Time step can be \sqrt{3}=1.73 bigger with 4th order (wavesolver.py:41-46)
        if self.kernel == 'OT4':
            return self.model.dtype(1.73 * self.model.critical_dt)
        return self.model.critical_dt

    # ------------------------------------------------------------------ helpers
    def _gs(self):
        return grid_struct(self.model.grid, self.space_order, self.kernel)

    def _field_dev(self, f, write=False):
        return f._buf.dev(write=write)

    def _vp_dev(self, vp):
        """Device slice of the velocity: Function, Constant or python scalar (acoustic_example.py:44-47)."""
        import torch
        if isinstance(vp, Function):
            return vp._buf.dev()
        value = float(vp.data if isinstance(vp, Constant) else vp)
        grid = self.model.grid
        t = torch.zeros(grid.slice_shape, dtype=torch.float32, device='cuda')
        t[tuple(slice(HALO, HALO + n) for n in grid.shape)] = value
        return t

    def _damp_dev(self):
        import torch
        damp = self.model.damp
        if isinstance(damp, Function):
            return damp._buf.dev()
        return torch.zeros(self.model.grid.slice_shape, dtype=torch.float32, device='cuda')   # nbl == 0

    def _coeffs(self, vp_dev, dt):
        import torch
        lib = _lib.lib()
        # two coefficient slices + B2FWI_COEF_TAIL floats (interior box written by the library)
        coef = torch.empty(2 * self.model.grid.slice_elems + 8, dtype=torch.float32, device='cuda')
        g = self._gs()
        _lib.check(lib.b2fwi_prepare_coeffs(ctypes.byref(g), _ptr(vp_dev), _ptr(self._damp_dev()),
                                            ctypes.c_float(dt), _ptr(coef), _stream()))
        return coef

    @staticmethod
    def _time_bounds(kwargs, nt):
        time_m = int(kwargs.pop('time_m', 1))
        time_M = int(kwargs.pop('time_M', nt - 2))
        return time_m, time_M

    def _summary(self, name, seconds, steps, bpp):
        if not self._profile:
            return None
        R = self.space_order // 2
        flops = 2 * self.model.dim * (2 * R) + 6
        return PerformanceSummary(name, seconds, np.prod(self.model.grid.shape), steps, bpp, flops)

    # ------------------------------------------------------------------ operators
    def forward(self, src=None, rec=None, u=None, vp=None, save=None, **kwargs):
        """Forward modelling: returns (rec, u, summary)   [wavesolver.py:76-114].
        Extension: ``save='checkpoint'`` keeps on-device checkpoints instead of the full history and
        returns a CheckpointedWavefield as ``u`` (accepted by ``gradient``), see checkpoint.py."""
        lib = _lib.lib()
        src = src or self.geometry.src
        rec = rec or self.geometry.rec
        if isinstance(save, str):
            if save != 'checkpoint':
                raise ValueError("save must be None, a bool or 'checkpoint'")
            from .checkpoint import checkpointed_forward
            dt = float(kwargs.pop('dt', self.dt))
            illum = kwargs.pop('illum', None)
            timer = _Timer(self._profile)
            cw = checkpointed_forward(self, src, rec, vp or self.model.vp, dt, illum=illum, **kwargs)
            summary = self._summary('Forward', timer.stop(), max(cw.time_M - cw.time_m + 1, 0), BYTES_FWD)
            return rec, cw, summary
        u = u or TimeFunction(name='u', grid=self.model.grid,
                              save=self.geometry.nt if save else None,
                              time_order=2, space_order=self.space_order)
        vp = vp or self.model.vp
        dt = float(kwargs.pop('dt', self.dt))
        illum = kwargs.pop('illum', None)            # extension: Function accumulating sum_t u^2
        nt = min(src.nt, rec.nt, u.save) if u.save else min(src.nt, rec.nt)
        time_m, time_M = self._time_bounds(kwargs, nt)

        grid = self.model.grid
        vp_dev = self._vp_dev(vp)
        coef = self._coeffs(vp_dev, dt)
        src_map = sparse_map(grid, src.coordinates.data)
        rec_map = sparse_map(grid, rec.coordinates.data)
        src_dev = src._sdata.dev()
        rec_dev = rec._sdata.dev(write=True)
        u_dev = self._field_dev(u, write=True)
        illum_dev = illum._buf.dev(write=True) if illum is not None else None
        g = self._gs()
        timer = _Timer(self._profile)
        _lib.check(lib.b2fwi_forward(
            ctypes.byref(g), _ptr(vp_dev), _ptr(coef), ctypes.c_float(dt), nt, time_m, time_M,
            _ptr(src_dev), src_map.byref(), _ptr(rec_dev), rec_map.byref(),
            _ptr(u_dev), 1 if u.save else 0, _ptr(illum_dev), None, 0, _stream()))
        summary = self._summary('Forward', timer.stop(), max(time_M - time_m + 1, 0), BYTES_FWD)
        return rec, u, summary

    def adjoint(self, rec, srca=None, v=None, vp=None, **kwargs):
        """Adjoint modelling: returns (srca, v, summary)   [wavesolver.py:116-151]."""
        lib = _lib.lib()
        srca = srca or self.geometry.new_src(name='srca', src_type=None)
        v = v or TimeFunction(name='v', grid=self.model.grid,
                              time_order=2, space_order=self.space_order)
        vp = vp or self.model.vp
        dt = float(kwargs.pop('dt', self.dt))
        nt = min(srca.nt, rec.nt)
        time_m, time_M = self._time_bounds(kwargs, nt)

        grid = self.model.grid
        vp_dev = self._vp_dev(vp)
        coef = self._coeffs(vp_dev, dt)
        rec_map = sparse_map(grid, rec.coordinates.data)
        src_map = sparse_map(grid, srca.coordinates.data)
        g = self._gs()
        rec_dev = rec._sdata.dev()
        srca_dev = srca._sdata.dev(write=True)
        v_dev = self._field_dev(v, write=True)
        timer = _Timer(self._profile)
        _lib.check(lib.b2fwi_adjoint(
            ctypes.byref(g), _ptr(vp_dev), _ptr(coef), ctypes.c_float(dt), nt, time_m, time_M,
            _ptr(rec_dev), rec_map.byref(), _ptr(srca_dev), src_map.byref(), _ptr(v_dev), _stream()))
        summary = self._summary('Adjoint', timer.stop(), max(time_M - time_m + 1, 0), BYTES_FWD)
        return srca, v, summary

    def jacobian_adjoint(self, rec, u, v=None, grad=None, vp=None,
                         checkpointing=False, **kwargs):
        """Gradient (adjoint Jacobian applied to ``rec``): returns (grad, summary)
        [wavesolver.py:153-205].  ``checkpointing=True`` re-computes the forward wavefield from
        on-device checkpoints instead of reading ``u`` (the pyrevolve branch, :188-201)."""
        lib = _lib.lib()
        dt = float(kwargs.pop('dt', self.dt))
        grad = grad or Function(name='grad', grid=self.model.grid)
        v_from_rest = v is None            # a caller-supplied v is an initial adjoint state (ownership)
        v = v or TimeFunction(name='v', grid=self.model.grid,
                              time_order=2, space_order=self.space_order)
        vp = vp or self.model.vp
        from .checkpoint import checkpointed_gradient, CheckpointedWavefield
        if isinstance(u, CheckpointedWavefield):
            return checkpointed_gradient(self, rec, v, grad, vp, dt, checkpoints=u, v_from_rest=v_from_rest, **kwargs)
        if checkpointing:
            return checkpointed_gradient(self, rec, v, grad, vp, dt, v_from_rest=v_from_rest, **kwargs)

        if not u.save:
            raise ValueError("the gradient needs the saved forward wavefield: forward(save=True)")
        nt = min(rec.nt, u.save)
        time_m, time_M = self._time_bounds(kwargs, nt)
        grid = self.model.grid
        vp_dev = self._vp_dev(vp)
        coef = self._coeffs(vp_dev, dt)
        rec_map = sparse_map(grid, rec.coordinates.data)
        g = self._gs()
        rec_dev = rec._sdata.dev()
        u_dev = self._field_dev(u)
        v_dev = self._field_dev(v, write=True)
        grad_dev = self._field_dev(grad, write=True)
        timer = _Timer(self._profile)
        _lib.check(lib.b2fwi_gradient(
            ctypes.byref(g), _ptr(vp_dev), _ptr(coef), ctypes.c_float(dt), nt, time_m, time_M,
            _ptr(rec_dev), rec_map.byref(), _ptr(u_dev), 1, 0, _ptr(v_dev), _ptr(grad_dev), _stream()))
        summary = self._summary('Gradient', timer.stop(), max(time_M - time_m + 1, 0), BYTES_ADJ)
        return grad, summary

    def jacobian(self, dmin, src=None, rec=None, u=None, U=None, vp=None, **kwargs):
        """Linearised (Born) modelling: returns (rec, u, U, summary)   [wavesolver.py:207-242].
        ``dmin``: perturbation of the squared slowness, a Function or an array of the padded grid shape."""
        import torch
        lib = _lib.lib()
        src = src or self.geometry.src
        rec = rec or self.geometry.rec
        u = u or TimeFunction(name='u', grid=self.model.grid, time_order=2, space_order=self.space_order)
        U = U or TimeFunction(name='U', grid=self.model.grid, time_order=2, space_order=self.space_order)
        if u.save or U.save:
            raise ValueError("the Born operator runs on ring-buffer wavefields (save=None), as in the reference")
        vp = vp or self.model.vp
        dt = float(kwargs.pop('dt', self.dt))
        nt = min(src.nt, rec.nt)
        time_m, time_M = self._time_bounds(kwargs, nt)
        grid = self.model.grid
        if isinstance(dmin, Function):
            dm = dmin
        else:
            dm = Function(name='dm', grid=grid, space_order=0)
            arr = np.asarray(dmin, dtype=np.float32)
            if arr.shape != grid.shape:
                raise ValueError("dm must have the padded grid shape %s (got %s)" % (grid.shape, arr.shape))
            dm.data[...] = arr
        vp_dev = self._vp_dev(vp)
        coef = self._coeffs(vp_dev, dt)
        src_map = sparse_map(grid, src.coordinates.data)
        rec_map = sparse_map(grid, rec.coordinates.data)
        src_dev = src._sdata.dev()
        rec_dev = rec._sdata.dev(write=True)
        u_dev, U_dev = self._field_dev(u, write=True), self._field_dev(U, write=True)
        scratch = torch.empty(grid.slice_shape, dtype=torch.float32, device='cuda')
        g = self._gs()
        timer = _Timer(self._profile)
        _lib.check(lib.b2fwi_born(
            ctypes.byref(g), _ptr(vp_dev), _ptr(coef), ctypes.c_float(dt), nt, time_m, time_M,
            _ptr(src_dev), src_map.byref(), _ptr(rec_dev), rec_map.byref(),
            _ptr(dm._buf.dev()), _ptr(u_dev), _ptr(U_dev), _ptr(scratch), _stream()))
        summary = self._summary('Born', timer.stop(), max(time_M - time_m + 1, 0), 2 * BYTES_FWD + 16)
        return rec, u, U, summary

    # Backward compatibility
    born = jacobian
    gradient = jacobian_adjoint
