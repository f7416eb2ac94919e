"""FWI objective ``(f, g)`` on the B200 propagator (API of the reference's fwi.py).

Same functions, arguments and return values as the reference so that ``minimize.py``, ``optimize/``
and ``misfit/`` run unchanged on top (fwi.py:59-81 fm_single/fm_multi, :104-129
fix_source_illumination, :131-173 fwi_obj_single, :175-205 fwi_obj_multi, :236-246 fwi_loss).
What moved: the per-shot host work (crop, 1+nrec Gaussian masks, the sum of u^2 over the whole
history) runs on the device; the illumination is accumulated by the forward kernel itself; and the
shot loop is partitioned over ranks with one all-reduce of [grad | illum | fval] (dist.py).
The misfit stays a host plug-in ``misfit_func(syn, obs) -> (fval, adjoint_source)`` on numpy arrays.
"""
import ctypes
import warnings
from copy import deepcopy

import numpy as np
from scipy import interpolate

from . import _lib, dist
from .grid import Function, TimeFunction
from .source import Receiver
from .geometry import AcquisitionGeometry
from .wavesolver import AcousticWaveSolver, grid_struct, _ptr, _stream

__all__ = ['fm_single', 'fm_multi', 'fwi_obj_single', 'fwi_obj_multi', 'fwi_loss',
           'fix_source_illumination', 'resample', 'Filter', 'least_square']


def least_square(x, y):
    """0.5*||x - y||^2 and its derivative (misfit/misfit.py:5-9)."""
    r = x - y
    return .5 * np.linalg.norm(r.flatten())**2, r


class Filter(object):
    """Butterworth band/low/high-pass applied to the source wavelet (fwi.py:31-44), via scipy."""

    def __init__(self, filter_type, freqmin=None, freqmax=None, df=None, corners=10,
                 zerophase=False, axis=-1):
        assert filter_type.lower() in ['bandpass', 'lowpass', 'highpass']
        self.filter_type = filter_type
        self.freqmin, self.freqmax, self.df = freqmin, freqmax, df
        self.corners, self.zerophase, self.axis = corners, zerophase, axis

    def __call__(self, data):
        """seismic_filter (fwi.py:10-29) -> seismic/filter/filter.py: bandpass (:33-73, falls back to a high-pass when
        the high corner reaches Nyquist), lowpass (:115-147, corner clamped to Nyquist), highpass (:150-182);
        zerophase runs the second pass on the record reversed along its FIRST axis, as the reference does."""
        import warnings
        from scipy.signal import iirfilter, sosfilt, zpk2sos
        ftype = self.filter_type
        if ftype == 'bandpass':
            if not (self.freqmin and self.freqmax and self.df):
                raise ValueError
        elif ftype == 'lowpass':
            if not (self.freqmax and self.df):
                raise ValueError
        elif not (self.freqmin and self.df):
            raise ValueError
        fe = 0.5 * self.df
        axis = self.axis
        if ftype == 'bandpass':
            low, high = self.freqmin / fe, self.freqmax / fe
            if high - 1.0 > -1e-6:
                warnings.warn("Selected high corner frequency ({}) of bandpass is at or above Nyquist ({}). "
                              "Applying a high-pass instead.".format(self.freqmax, fe))
                ftype, axis = 'highpass', -1          # the reference's fall-back call does not pass `axis` on
            elif low > 1:
                raise ValueError("Selected low corner frequency is above Nyquist.")
            else:
                wn, btype = [low, high], 'band'
        if ftype == 'lowpass':
            f = self.freqmax / fe
            if f > 1:
                f = 1.0
                warnings.warn("Selected corner frequency is above Nyquist. Setting Nyquist as high corner.")
            wn, btype = f, 'lowpass'
        elif ftype == 'highpass':
            f = self.freqmin / fe
            if f > 1:
                raise ValueError("Selected corner frequency is above Nyquist.")
            wn, btype = f, 'highpass'
        z, p, k = iirfilter(self.corners, wn, btype=btype, ftype='butter', output='zpk')
        sos = zpk2sos(z, p, k)
        out = sosfilt(sos, data, axis)
        if self.zerophase:
            out = sosfilt(sos, out[::-1], axis)[::-1]
        return out

    def key(self):
        """Everything the filtered wavelet depends on (cache key of the resident surveys)."""
        return (self.filter_type, self.freqmin, self.freqmax, self.corners, self.zerophase, self.axis)


def resample(x, t, t0, order=3):
    """Cubic-spline resampling of traces from axis t0 to axis t; identity for equal steps (fwi.py:47-57)."""
    if np.isclose(t[1] - t[0], t0[1] - t0[0]):
        return x
    out = np.zeros((t.size, x.shape[1]), dtype=np.float32)
    for i in range(x.shape[1]):
        out[:, i] = interpolate.splev(t, interpolate.splrep(t0, x[:, i], k=order))
    return out


def _shot_geometry(geometry, i):
    return AcquisitionGeometry(geometry.model, geometry.rec_positions, geometry.src_positions[i, :],
                               geometry.t0, geometry.tn, f0=geometry.f0, src_type=geometry.src_type,
                               filter=geometry._filter)


def fm_single(geometry, save=False):
    """Forward modelling of one shot: (Receiver, wavefield)   [fwi.py:59-65]."""
    solver = AcousticWaveSolver(geometry.model, geometry, space_order=geometry.model.space_order,
                                profile=False)
    data, u = solver.forward(vp=geometry.model.vp, save=save)[0:2]
    return data, u


def fm_multi(geometry, save=False):
    """Forward modelling of every shot of a survey: list of Receivers   [fwi.py:67-81].
    2-D surveys run as one batched launch of the SM-resident engine (all shots concurrently)."""
    surveys = _resident_surveys(geometry, list(range(geometry.nsrc))) if not save else None
    if not surveys:
        return [fm_single(_shot_geometry(geometry, i), save)[0] for i in range(geometry.nsrc)]
    shots = []
    for survey in surveys:
        rec = survey.forward(save=False).clone()
        for k in range(survey.nshots):
            r = Receiver(name='rec', grid=geometry.grid, time_range=geometry.time_axis,
                         coordinates=geometry.rec_positions)
            r._sdata.adopt_dev(rec[k])
            shots.append(r)
    return shots


def fix_source_illumination(geometry, g):
    """Host version of the source/receiver muting (fwi.py:104-129), axis swap included; the
    objective functions use the device kernel b2fwi_geometry_mask instead."""
    if geometry.src_positions.shape[0] > 1:
        raise ValueError("Only single source valid.")
    dx, dz = geometry.model.spacing
    nx, nz = geometry.model.shape
    if g.shape != (nx, nz):
        raise ValueError("Shape does not match!")
    xx, zz = np.meshgrid(np.arange(0, nz) * dz, np.arange(0, nx) * dx)
    sigma = dx + dz
    pts = np.concatenate([geometry.src_positions[:1], geometry.rec_positions], axis=0)
    for c0, c1 in pts:
        g = g * (1. - np.exp(-.5 * ((xx - c0)**2 + (zz - c1)**2) / (sigma**2)))
    return g


# ---------------------------------------------------------------------------------------------
_WORKSPACE = {}


CHECKPOINT = None    # None: checkpoint when the saved history would not fit; True / False: force (tests)


def _use_checkpoints(model, nt):
    if CHECKPOINT is not None:
        return bool(CHECKPOINT)
    import torch
    hist = float(nt) * model.grid.slice_elems * 4
    key = ('u', model.grid._key(), nt)
    if key in _WORKSPACE:            # the history buffer already exists
        return False
    return hist > 0.6 * torch.cuda.mem_get_info()[0]


def _saved_wavefield(model, nt, space_order):
    """Re-used nt-slice history buffer: only slots 0 and 1 need zeroing (every other slot is
    overwritten by the forward sweep), instead of a fresh zero-filled TimeFunction per shot."""
    key = ('u', model.grid._key(), nt)
    u = _WORKSPACE.get(key)
    if u is None:
        _WORKSPACE.clear()
        u = _WORKSPACE[key] = TimeFunction(name='u', grid=model.grid, save=nt, time_order=2,
                                           space_order=space_order)
    else:
        d = u._buf.dev(write=True)
        d[0:2].zero_()
    return u


_MASKS = {}


def _geometry_mask(geometry):
    """Device fp64 mask prod_k(1 - G_k) of one shot geometry, cached by content."""
    import torch
    model = geometry.model
    pts = np.ascontiguousarray(np.concatenate([geometry.src_positions[:1], geometry.rec_positions]),
                               dtype=np.float64)
    key = (model.grid._key(), model.nbl, pts.tobytes())
    m = _MASKS.get(key)
    if m is None:
        if len(_MASKS) > 512:
            _MASKS.clear()
        g = grid_struct(model.grid, model.space_order)
        pts_dev = torch.from_numpy(pts).cuda()
        m = torch.empty(model.shape, dtype=torch.float64, device='cuda')
        _lib.check(_lib.lib().b2fwi_geometry_mask(ctypes.byref(g), model.nbl, _ptr(pts_dev), pts.shape[0],
                                                  _ptr(m), _stream()))
        _MASKS[key] = m
    return m


def _crop_mask_acc(geometry, field_dev, mask, out):
    model = geometry.model
    g = grid_struct(model.grid, model.space_order)
    _lib.check(_lib.lib().b2fwi_crop_mask_accumulate(ctypes.byref(g), model.nbl, _ptr(field_dev), _ptr(mask),
                                                     _ptr(out), _stream()))


def _fwi_obj_single_dev(geometry, obs, misfit_func, direct_wave, resample_dt, calc_grad, acc):
    """One shot; accumulates the masked cropped gradient / illumination into ``acc`` (device fp64,
    [2, nx, nz]) and returns (fval, residual ndarray)."""
    if geometry.src_positions.shape[0] > 1:
        raise ValueError("Only single source valid.")
    model = geometry.model
    solver = AcousticWaveSolver(model, geometry, space_order=model.space_order, profile=False)
    illum = Function(name='illum', grid=model.grid) if calc_grad else None
    if calc_grad and _use_checkpoints(model, geometry.nt):
        # 3-D: the history does not fit in HBM -> on-device checkpoints, recomputed by the gradient (checkpoint.py)
        _WORKSPACE.clear()
        pred, wfd = solver.forward(vp=model.vp, save='checkpoint', illum=illum)[0:2]
    else:
        wfd = _saved_wavefield(model, geometry.nt, model.space_order) if calc_grad else None
        pred, wfd = solver.forward(vp=model.vp, save=calc_grad, u=wfd, illum=illum)[0:2]

    dw = direct_wave
    if resample_dt is None:
        resample_dt = geometry.dt
    else:
        obs = deepcopy(obs).resample(resample_dt) if not np.isclose(resample_dt, obs.time_range.step) else obs
        pred = pred.resample(resample_dt)
        if direct_wave is not None:
            dw = direct_wave if np.isclose(resample_dt, direct_wave.time_range.step) \
                else deepcopy(direct_wave).resample(resample_dt)
    syn_data = pred.data
    obs_data = obs.data
    if direct_wave is not None:
        syn_data = syn_data - dw.data
        obs_data = obs_data - dw.data
    fval, residual_data = misfit_func(syn_data, obs_data)

    residual = Receiver(name="rec", grid=model.grid, time_range=geometry.time_axis,
                        coordinates=geometry.rec_positions)
    residual.data[:] = resample(residual_data, geometry.time_axis.time_values, pred.time_values)[:]
    if calc_grad:
        grad = Function(name="grad", grid=model.grid)
        solver.gradient(rec=residual, u=wfd, vp=model.vp, grad=grad)
        mask = _geometry_mask(geometry)
        _crop_mask_acc(geometry, grad._buf.dev(), mask, acc[0])
        _crop_mask_acc(geometry, illum._buf.dev(), mask, acc[1])
    return fval, residual.data


def _filter_key(flt):
    if flt is None:
        return None
    return flt.key() if hasattr(flt, 'key') else tuple(sorted((k, repr(v)) for k, v in vars(flt).items() if k != 'df'))


_SURVEYS = {}
ENGINE = 'auto'     # 'auto' | 'stream' (force the per-shot streaming engine; used by the parity tests)


_FALLBACK_WARNED = set()


def _warn_streaming_fallback(geometry):
    """A 2-D survey the SM-resident engine cannot take (grid beyond 16 SMs' shared memory, space order > 8, free
    surface, ...) runs shot by shot on the streaming engine, one launch per time step: correct, but an order of
    magnitude slower per shot on grids this small - say so once per grid instead of dropping there silently."""
    model = geometry.model
    if model.grid.dim != 2:
        return
    key = (model.grid._key(), model.space_order, bool(getattr(model, 'fs', False)))
    if key not in _FALLBACK_WARNED:
        _FALLBACK_WARNED.add(key)
        warnings.warn("2-D model %s (space_order %d%s) does not fit the SM-resident engine: running shot by shot on the "
                      "streaming engine (one launch per time step)" % (tuple(model.grid.shape), model.space_order,
                                                                      ", free surface" if key[2] else ""))


def _resident_surveys(geometry, shots):
    """Cached list of ResidentSurvey objects covering ``shots`` in order (one launch group each, see
    resident.partition_shots; a single group whenever all shots fit one wave of clusters), or None when the
    SM-resident engine does not apply."""
    from .resident import ResidentSurvey, partition_shots
    if ENGINE == 'stream' or not shots:
        return None
    if not ResidentSurvey.supported(geometry):
        _warn_streaming_fallback(geometry)
        return None
    model = geometry.model
    key = (id(model), model.grid._key(), model.space_order, tuple(shots), float(geometry.dt), geometry.nt,
           geometry.src_positions.tobytes(), geometry.rec_positions.tobytes(), geometry.f0, geometry.src_type,
           _filter_key(geometry._filter))
    svs = _SURVEYS.get(key)
    if svs is None:
        if len(_SURVEYS) >= 4:
            _SURVEYS.clear()
        groups = partition_shots(model.grid, model.space_order, model.nbl, len(shots))
        if not groups:
            return None
        # the u.dt2 history of every launch group stays allocated with the cached survey: when it would not fit in what
        # is free now, use the streaming engine (per-shot history / checkpoints) instead of running into an OOM
        import torch
        p0 = groups[0][1]
        per_shot = (geometry.nt - 2) * (p0.wx1 - p0.wx0) * (p0.wq1 - p0.wq0) * 16 + 2 * geometry.nt * geometry.nrec * 4
        cached = sum(sv.nbytes for key_ in _SURVEYS for sv in _SURVEYS[key_])
        if per_shot * len(shots) > 0.85 * (torch.cuda.mem_get_info()[0] + cached):
            return None
        if per_shot * len(shots) > 0.85 * torch.cuda.mem_get_info()[0]:
            _SURVEYS.clear()              # idle surveys of other geometries hold the memory: evict them
            torch.cuda.empty_cache()
        try:
            svs, k0 = [], 0
            for count, plan in groups:
                svs.append(ResidentSurvey(geometry, shots[k0:k0 + count], plan=plan))
                k0 += count
        except ValueError:
            return None
        _SURVEYS[key] = svs
    return svs


def _resident_survey(geometry, shots):
    """The survey of ``shots`` when they form a single launch group (the common case), else None."""
    svs = _resident_surveys(geometry, shots)
    return svs[0] if svs and len(svs) == 1 else None


class LazyResidual(object):
    """Residual of one shot living on the device (a private snapshot of this evaluation's adjoint source); converts
    to numpy on first host use and then behaves like the ndarray the reference returns (minimize.py only dumps the
    residuals every few iterations, minimize.py:50-51,146-152)."""

    def __init__(self, tensor):
        self._t = tensor
        self._h = None
        self.shape = tuple(tensor.shape)
        self.dtype = np.dtype(np.float32)
        self.ndim = len(self.shape)
        self.size = int(np.prod(self.shape))

    def __array__(self, dtype=None, copy=None):
        if self._h is None:
            self._h = self._t.cpu().numpy()
            self._t = None
        return self._h if dtype is None else self._h.astype(dtype)

    def __getattr__(self, name):          # .copy, .T, .ravel, .sum, .tofile, ... : whatever an ndarray has
        if name.startswith('__'):
            raise AttributeError(name)
        return getattr(np.asarray(self), name)

    def __getitem__(self, idx):
        return np.asarray(self)[idx]

    def __len__(self):
        return self.shape[0]

    def __iter__(self):
        return iter(np.asarray(self))


def _lazy_binop(name):
    def op(self, *args):
        return getattr(np.asarray(self), name)(*args)
    op.__name__ = name
    return op


for _n in ('add', 'sub', 'mul', 'truediv', 'floordiv', 'pow', 'matmul', 'radd', 'rsub', 'rmul', 'rtruediv', 'rpow',
           'rmatmul', 'neg', 'pos', 'abs', 'lt', 'le', 'gt', 'ge', 'eq', 'ne'):
    setattr(LazyResidual, '__%s__' % _n, _lazy_binop('__%s__' % _n))
LazyResidual.__hash__ = None


def _from_misfit_module(obj):
    """True for objects defined in this module or in the reference's misfit/misfit.py (module `misfit.misfit`, also
    reached as devito_fwi_b200.compat...): a user's own callable that merely shares a name is NOT replaced."""
    mod = getattr(obj, '__module__', '') or ''
    return mod == __name__ or mod.split('.')[-1] == 'misfit'


def _host_forced(misfit_func):
    return bool(getattr(misfit_func, 'b2fwi_host', False))        # explicit opt-out: run as a host plug-in


def _is_l2(misfit_func):
    """The reference's least_square (misfit/misfit.py:5-9) or ours."""
    if _host_forced(misfit_func):
        return False
    return misfit_func is least_square or (getattr(misfit_func, '__name__', '') == 'least_square' and
                                           _from_misfit_module(misfit_func))


def _is_qw(misfit_func, method):
    cls = type(misfit_func)
    return (not _host_forced(misfit_func) and cls.__name__ == 'qWasserstein' and _from_misfit_module(cls) and
            getattr(misfit_func, 'method', None) == method and getattr(misfit_func, 'trans_type', None) == 'linear')


def _is_w1d(misfit_func):
    """The reference's qWasserstein(trans_type='linear', method='1d') instance (misfit/misfit.py:11-104)."""
    return _is_qw(misfit_func, '1d')


def _is_w2d(misfit_func):
    """qWasserstein(trans_type='linear', method='2d'): the back-and-forth solver misfit/QW2D (bfm.py:145-193)."""
    return _is_qw(misfit_func, '2d')


def _bfm_params(misfit_func):
    """(num_steps, step_scale) of a qWasserstein object: the reference keeps them on its `bfm` solver
    (misfit/misfit.py:17, bfm.py:149-154), devito_fwi_b200.misfit.qWasserstein on itself."""
    holder = getattr(misfit_func, 'bfm', None) or misfit_func
    return int(getattr(holder, 'num_steps')), float(getattr(holder, 'step_scale'))


def _stack_dev(receivers, shots, cache_owner, tag, stream=None, after=None):
    """[nshots, nt, nrec] device tensor of a list of Receivers; re-used while the SAME list object is
    passed again and no record's host view has been handed out since (``.data`` access = possibly new
    host data = upload again). With ``stream`` the uploads are issued on that (copy) stream from the
    records' pinned host buffers, so they overlap the forward sweep; ``after``: event marking the last
    device-side read of the previous contents (the copy stream waits for it instead of for everything
    queued on the compute stream - the forward sweep has already been launched there)."""
    import torch
    cache = cache_owner.__dict__.setdefault('_dev_stacks', {})
    hit = cache.get(tag)
    versions = tuple(receivers[i]._sdata._hver for i in shots)
    if hit is not None and hit[0] is receivers and hit[1] == tuple(shots) and hit[3] == versions:
        return hit[2]
    sd0 = receivers[shots[0]]._sdata
    if hit is not None and tuple(hit[2].shape) == (len(shots),) + sd0.shape:
        t = hit[2]
    else:
        t = torch.empty((len(shots),) + sd0.shape, dtype=torch.float32, device='cuda')
    on_dev = [receivers[i]._sdata._newer == 'dev' or receivers[i]._sdata._host is None for i in shots]
    if stream is not None:
        if after is not None and not any(on_dev):
            stream.wait_event(after)
        else:       # device-resident records were produced on the compute stream: order the copies behind it
            stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream if stream is not None else torch.cuda.current_stream()):
        for k, i in enumerate(shots):
            sd = receivers[i]._sdata
            if on_dev[k]:
                t[k].copy_(sd.dev())                                   # record lives on the device
            else:
                t[k].copy_(sd._host_t if sd._host_t is not None else torch.from_numpy(sd._host),
                           non_blocking=True)
    cache[tag] = (receivers, tuple(shots), t, versions)
    return t


def _objective_resident(survey, geometry, obs, misfit_func, direct_wave, calc_grad, acc):
    """Shots of ``survey`` in two launches (forward, backward); returns (fval, residual list)."""
    import torch
    lib = _lib.lib()
    shots = survey.shots
    w1d, w2d = _is_w1d(misfit_func), _is_w2d(misfit_func)
    l2 = _is_l2(misfit_func) or w1d or w2d   # misfits evaluated on the device
    if l2 and getattr(survey, '_copy_stream', None) is None:
        survey._copy_stream = torch.cuda.Stream()
        survey._misfit_done = None
        # first use: the stacks are allocated before the sweep is queued
        _stack_dev(obs, shots, survey, 'obs', survey._copy_stream)
        if direct_wave is not None:
            _stack_dev(direct_wave, shots, survey, 'dw', survey._copy_stream)
    syn = survey.forward(save=calc_grad, illum=calc_grad)       # queued first: the GPU starts right away
    if l2:
        # observed / direct-wave records go up on a copy stream while the forward sweep runs
        done = survey._misfit_done
        obs_d = _stack_dev(obs, shots, survey, 'obs', survey._copy_stream, done)
        dw_d = _stack_dev(direct_wave, shots, survey, 'dw', survey._copy_stream, done) \
            if direct_wave is not None else None
        # on-device least squares (misfit/misfit.py:5-9) incl. direct-wave subtraction (fwi.py:146-150)
        torch.cuda.current_stream().wait_stream(survey._copy_stream)
        if getattr(survey, '_res', None) is None:
            survey._res = torch.empty_like(syn)
            survey._fval = torch.zeros(1, dtype=torch.float64, device='cuda')
            survey._scratch = torch.empty(1024, dtype=torch.float64, device='cuda')
        survey._fval.zero_()
        if w1d:
            ns, nt, nrec = syn.shape
            if getattr(survey, '_w1d_scratch', None) is None:
                nbytes = int(lib.b2fwi_w1d_scratch_bytes(nt, nrec, ns))
                survey._w1d_scratch = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
            _lib.check(lib.b2fwi_w1d_misfit(_ptr(syn), _ptr(obs_d), _ptr(dw_d), nt, nrec, ns,
                                            ctypes.c_double(float(misfit_func.gamma)), _ptr(survey._res),
                                            _ptr(survey._fval), _ptr(survey._w1d_scratch), _stream()))
        elif w2d:
            ns, nt, nrec = syn.shape
            if getattr(survey, '_w2d_scratch', None) is None:
                nbytes = int(lib.b2fwi_qw2d_scratch_bytes(nt, nrec, ns))
                survey._w2d_scratch = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
            num_steps, step_scale = _bfm_params(misfit_func)
            _lib.check(lib.b2fwi_qw2d_misfit(_ptr(syn), _ptr(obs_d), _ptr(dw_d), nt, nrec, ns,
                                             ctypes.c_double(float(misfit_func.gamma)), num_steps,
                                             ctypes.c_float(step_scale), _ptr(survey._res), _ptr(survey._fval), None,
                                             _ptr(survey._w2d_scratch), _stream()))
        else:
            _lib.check(lib.b2fwi_l2_misfit(_ptr(syn), _ptr(obs_d), _ptr(dw_d), syn.numel(), _ptr(survey._res),
                                           _ptr(survey._fval), _ptr(survey._scratch), _stream()))
        survey._misfit_done = torch.cuda.Event()
        survey._misfit_done.record()
        residual = survey._res
        fval = survey._fval            # stays on the device until the all-reduce
        # the caller's residuals are a private snapshot: survey._res is overwritten by the next evaluation (a line-search
        # trial), the reference returns independent arrays (47 MB device copy for 29 Marmousi shots, ~20 us)
        snap = residual.clone()
        residuals = [LazyResidual(snap[k]) for k in range(len(shots))]
    else:
        # host plug-in misfit: misfit_func(syn, obs) -> (fval, adjoint_source) on numpy arrays
        syn_h = syn.cpu().numpy()
        res_h = np.empty_like(syn_h)
        fval = 0.
        for k, i in enumerate(shots):
            syn_data, obs_data = syn_h[k], obs[i].data
            if direct_wave is not None:
                syn_data = syn_data - direct_wave[i].data
                obs_data = obs_data - direct_wave[i].data
            f_, r_ = misfit_func(syn_data, obs_data)
            fval += f_
            res_h[k] = r_
        residual = torch.from_numpy(res_h).cuda()
        residuals = [res_h[k] for k in range(len(shots))]
    if calc_grad:
        grad = survey.gradient(residual)
        survey.window_mask_accumulate_all(grad, acc[0])
        survey.window_mask_accumulate_all(survey.illum, acc[1])
    return fval, residuals


def fwi_obj_single(geometry, obs, misfit_func, direct_wave=None, resample_dt=None, calc_grad=False):
    """Objective and gradient of one shot: (fval, crop_grad, residual, illum)   [fwi.py:131-173]."""
    import torch
    acc = torch.zeros((2,) + tuple(geometry.model.shape), dtype=torch.float64, device='cuda')
    fval, res = _fwi_obj_single_dev(geometry, obs, misfit_func, direct_wave, resample_dt, calc_grad, acc)
    if not calc_grad:
        return fval, None, res, None
    host = acc.cpu().numpy()
    return fval, host[0], res, host[1]


def fwi_obj_multi(geometry, obs, misfit_func, direct_wave=None, mask=None, precond=True,
                  calc_grad=False):
    """Sum over shots: (fval, grad float64[nx*nz], residuals)   [fwi.py:175-205].

    Under torch.distributed the shots are split round-robin over ranks and [grad | illum | fval] is
    all-reduced once; every rank returns the same (fval, grad). ``residuals`` then holds the local
    shots' residuals only (they are only ever dumped to disk, minimize.py:50-51)."""
    import torch
    model = geometry.model
    n = int(np.prod(model.shape))
    buf = torch.zeros(2 * n + 1, dtype=torch.float64, device='cuda')
    acc = buf[:2 * n].view((2,) + tuple(model.shape))
    fval = .0
    residuals = []
    shots = dist.local_shots(geometry.nsrc)
    dt0 = float(geometry.dt)      # np.isclose semantics on plain floats (29 numpy calls cost 0.5 ms per evaluation)
    same_dt = all(abs(dt0 - float(obs[i].time_range.step)) <= 1e-8 + 1e-5 * abs(float(obs[i].time_range.step))
                  for i in shots)
    surveys = _resident_surveys(geometry, shots) if same_dt else None
    if surveys:
        import torch
        try:
            for survey in surveys:
                fval_, res_ = _objective_resident(survey, geometry, obs, misfit_func, direct_wave, calc_grad, acc)
                fval = fval_ + fval          # device tensor (on-device misfits) or float
                residuals += res_
        except torch.cuda.OutOfMemoryError:
            # history buffers did not fit after all (another tenant of the GPU, fragmentation): drop every cached
            # survey and redo this evaluation on the streaming engine
            import warnings
            warnings.warn("resident engine ran out of HBM; falling back to the streaming engine for this survey")
            _SURVEYS.clear()
            surveys, fval, residuals = None, .0, []
            buf.zero_()
            torch.cuda.empty_cache()
    if not surveys:
        for i in shots:
            geom_i = _shot_geometry(geometry, i)
            dw = direct_wave[i] if direct_wave is not None else None
            fval_, res_ = _fwi_obj_single_dev(geom_i, obs[i], misfit_func, dw, geometry.dt, calc_grad, acc)
            fval += fval_
            residuals += [res_]
    if hasattr(fval, 'is_cuda'):
        buf[2 * n:2 * n + 1].copy_(fval)
    else:
        buf[2 * n] = float(fval)
    if not calc_grad:
        # line-search evaluation (minimize.py:59-86): only fval crosses NVLink and PCIe
        tail = buf[2 * n:]
        dist.all_reduce_sum(tail)
        return float(tail.cpu().numpy()[0]), np.zeros(n, dtype=np.float64), residuals
    dist.all_reduce_sum(buf)
    return _finalize_objective(buf.cpu().numpy(), model.shape, mask, precond, calc_grad) + (residuals,)


def _finalize_objective(host, shape, mask, precond, calc_grad):
    """[grad | illum | fval] summed over shots and ranks -> (fval, grad): illumination preconditioning
    and bathymetry mask of fwi.py:200-204, applied identically (and redundantly) on every rank."""
    n = int(np.prod(shape))
    fval = float(host[2 * n])
    grad = host[:n].reshape(shape).copy()
    if calc_grad:
        if precond:
            grad /= np.sqrt(host[n:2 * n].reshape(shape) + 1e-30)
        if mask is not None:
            grad *= mask
    return fval, grad.reshape(-1).astype(np.float64)


def fwi_loss(x, geometry, obs, misfit_func, direct_wave=None, mask=None, precond=True, calc_grad=True):
    """Objective in squared slowness x = 1/vp^2 (flattened): (fval, grad, residuals)   [fwi.py:236-246]."""
    v = 1. / np.sqrt(x.reshape(geometry.model.shape))
    geometry.model.update('vp', v.reshape(geometry.model.shape))
    return fwi_obj_multi(geometry, obs, misfit_func, direct_wave, mask, precond, calc_grad)


# ---------------------------------------------------------------------------------------------
class StreamingSurvey(object):
    """L2 objective and gradient of a multi-shot survey on the STREAMING engine (any model the resident 2-D
    engine does not take: every 3-D model): the shot loop of fwi.py:183-199 with the per-shot sequence of
    acoustic_example.py:26-63 -- forward with on-device checkpoints, residual, adjoint + imaging (checkpoint.py)
    -- shots round-robin over ranks and ONE all-reduce of [grad | fval].
    No illumination preconditioning / source muting: fix_source_illumination is 2-D only (fwi.py:104-129)."""

    def __init__(self, geometry, keep_segments=None):
        import torch
        self.geometry = geometry
        self.model = geometry.model
        self.keep_segments = keep_segments
        grid = self.model.grid
        self.buf = torch.zeros(grid.slice_elems + 1, dtype=torch.float32, device='cuda')      # [grad | fval]
        self.grad = Function(name='grad', grid=grid)
        self.grad._buf._dev = self.buf[:-1].view(grid.slice_shape)       # accumulate straight into the all-reduce buffer
        self.grad._buf._newer = 'dev'
        self._host = None
        self.shots = dist.local_shots(geometry.nsrc)
        self._geoms = {i: _shot_geometry(geometry, i) for i in self.shots}
        self._solvers = {i: AcousticWaveSolver(self.model, g, space_order=self.model.space_order, profile=False)
                         for i, g in self._geoms.items()}

    def host_buffer(self):
        """Pinned host mirror of [grad | fval] (allocated on first use: pinning 0.85 GB takes ~0.4 s)."""
        import torch
        if self._host is None:
            self._host = torch.empty(self.buf.shape, dtype=torch.float32).pin_memory()
        return self._host.numpy()

    def forward(self, vp=None):
        """Synthetic records of this rank's shots (dict shot -> Receiver, data on the device)."""
        out = {}
        for i in self.shots:
            rec = self._solvers[i].forward(vp=vp or self.model.vp)[0]
            r = Receiver(name='obs', grid=self.model.grid, time_range=self._geoms[i].time_axis,
                         coordinates=self._geoms[i].rec_positions)
            r._sdata.adopt_dev(rec._sdata.dev().clone())
            out[i] = r
        return out

    def objective(self, obs, host=True):
        """(fval, grad): sum over all shots and ranks of 0.5*||d_syn - d_obs||^2 and its gradient with respect to
        the squared slowness on the padded grid. ``obs``: mapping shot -> Receiver. ``host=True`` returns
        (float, numpy [grid.shape]) through one device-to-host copy; ``host=False`` leaves both in the device
        buffer and returns (0-d tensor, Function)."""
        import torch
        self.buf.zero_()
        fval = torch.zeros(1, dtype=torch.float64, device='cuda')
        for i in self.shots:
            solver, g_i = self._solvers[i], self._geoms[i]
            rec, cw, _ = solver.forward(save='checkpoint', keep_segments=self.keep_segments)
            res = rec._sdata.dev() - obs[i]._sdata.dev()
            fval += 0.5 * (res.double() ** 2).sum()
            r = Receiver(name='res', grid=self.model.grid, time_range=g_i.time_axis, coordinates=g_i.rec_positions)
            r._sdata.adopt_dev(res)
            solver.gradient(rec=r, u=cw, grad=self.grad)
            del cw
        self.buf[-1:] = fval.float()
        dist.all_reduce_sum(self.buf)
        if not host:
            return self.buf[-1], self.grad
        h = self.host_buffer()
        self._host.copy_(self.buf)
        grid = self.model.grid
        g = h[:-1].reshape(grid.slice_shape)[tuple(slice(0, n) for n in grid.shape)]
        return float(h[-1]), g
