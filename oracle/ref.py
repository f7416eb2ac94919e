"""oracle/ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy/ctypes front end of the CPU restatement in ``fwi_oracle.c``: model set-up
(edge padding, damping profile, CFL time step, Ricker wavelet), the forward /
adjoint / gradient sweeps and the per-shot host post-processing of ``fwi.py``.
Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of
``bench.py`` may import this module; ``devito_fwi_b200`` never does.

Reference lines restated here (relative to /root/reference):
  seismic/model.py:13-51      initialize_damp            -> init_damp
  seismic/model.py:167-178    _gen_phys_param / padfunc  -> pad_edge
  seismic/model.py:338-370    _cfl_coeff / critical_dt   -> critical_dt
  seismic/source.py:42-75     TimeAxis                   -> time_axis
  seismic/source.py:272-277   RickerSource.wavelet       -> ricker
  fwi.py:104-129              fix_source_illumination
  fwi.py:131-205              fwi_obj_single / fwi_obj_multi
Parity status: pinned by tests/test_oracle_kat.py (see fwi_oracle.c header).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


class _Grid(ctypes.Structure):
    _fields_ = [("ndim", ctypes.c_int), ("shape", ctypes.c_int * 3),
                ("space_order", ctypes.c_int), ("spacing", ctypes.c_double * 3),
                ("origin", ctypes.c_double * 3), ("fs", ctypes.c_int), ("ot4", ctypes.c_int)]


def build(fast=False):
    target = "libfwi_oracle_fast.so" if fast else "libfwi_oracle.so"
    subprocess.run(["make", "-C", _HERE, target], check=True, stdout=subprocess.DEVNULL)
    return os.path.join(_HERE, target)


def lib(fast=False):
    if fast not in _LIBS:
        path = os.path.join(_HERE, "libfwi_oracle_fast.so" if fast else "libfwi_oracle.so")
        src_mtime = max(os.path.getmtime(os.path.join(_HERE, f))
                        for f in ("fwi_oracle.c", "fwi_oracle_body.inc"))
        if not os.path.exists(path) or os.path.getmtime(path) < src_mtime:
            build(fast)
        _LIBS[fast] = ctypes.CDLL(path)
    return _LIBS[fast]


# ----------------------------------------------------------------------------- set-up
def laplace_coeffs(space_order):
    c = (ctypes.c_double * 9)()
    lib().oracle_laplace_coeffs(int(space_order), c)
    return np.array(c[:space_order // 2 + 1])


def pad_edge(vp, nbl, fs=False):
    """Edge-replicated padding; with a free surface no layer above the last dimension (model.py:157)."""
    pads = [(nbl, nbl)] * vp.ndim
    if fs:
        pads[-1] = (0, nbl)
    return np.pad(vp, pads, mode="edge")


def init_damp(shape_padded, nbl, spacing, abc_type="damp", dtype=np.float32, fs=False):
    """model.py:13-51: additive per-dimension sponge profile (no left layer in the last dimension under fs, :33)."""
    damp = np.full(shape_padded, 1.0 if abc_type == "mask" else 0.0, dtype=np.float64)
    if nbl == 0:
        return damp.astype(dtype)
    dampcoeff = 1.5 * np.log(1.0 / 0.001) / nbl
    i = np.arange(nbl)                       # distance index from the outer edge
    pos = np.abs((nbl - i + 1) / float(nbl))
    val = dampcoeff * (pos - np.sin(2 * np.pi * pos) / (2 * np.pi))
    if abc_type == "mask":
        val = -val
    for d, h in enumerate(spacing):
        prof = np.zeros(shape_padded[d])
        if not (fs and d == len(spacing) - 1):
            prof[:nbl] += val / h
        prof[-nbl:] += (val / h)[::-1]
        sh = [1] * len(shape_padded)
        sh[d] = shape_padded[d]
        damp += prof.reshape(sh)
    return damp.astype(dtype)


def central_d2_weights(R):
    """Central 2nd-derivative weights on offsets -R..R (what sympy's fd_w(2, range(-R, R+1), 0) gives)."""
    from math import factorial
    k = np.arange(1, R + 1)
    ck = np.array([2.0 * (-1) ** (j + 1) * factorial(R) ** 2 /
                   (j * j * factorial(R - j) * factorial(R + j)) for j in k])
    return np.concatenate([ck[::-1], [-2.0 * ck.sum()], ck])


def cfl_coeff(space_order, ndim):
    """model.py:350-353 (acoustic branch; uses a 2*so-wide stencil)."""
    coeffs = central_d2_weights(space_order)
    return np.sqrt(4.0 / float(ndim * np.sum(np.abs(coeffs))))


def critical_dt(vp_max, spacing, space_order, dtype=np.float32, dt=None):
    """model.py:355-370."""
    cdt = cfl_coeff(space_order, len(spacing)) * np.min(spacing) / vp_max
    cdt = dtype("%.3e" % cdt)
    if dt:
        if dt > cdt:
            raise ValueError("Critical dt: %f, set dt: %f" % (cdt, dt))
        return dt
    return cdt


def time_axis(t0, tn, dt):
    """source.py:47-50: returns (num, stop, time_values)."""
    num = int(np.ceil((tn - t0 + dt) / dt))
    stop = dt * (num - 1) + t0
    return num, stop, np.linspace(t0, stop, num)


def ricker(f0, time_values, t0=None, a=None):
    """source.py:272-277."""
    t0 = t0 or 1.0 / f0
    a = a or 1
    r = np.pi * f0 * (time_values - t0)
    return a * (1 - 2. * r ** 2) * np.exp(-r ** 2)


class RefModel(object):
    """Numpy stand-in for seismic.Model (model.py:227-400), acoustic fields only."""

    def __init__(self, origin, spacing, shape, space_order, vp, nbl=20, dtype=np.float32,
                 dt=None, fs=False):
        self.shape = tuple(shape)
        self.spacing = tuple(float(s) for s in spacing)
        self.space_order = int(space_order)
        self.nbl = int(nbl)
        self.dtype = dtype
        self.origin = tuple(dtype(o) for o in origin)
        self.fs = bool(fs)
        origin_pml = [dtype(o - s * nbl) for o, s in zip(origin, spacing)]
        shape_pml = [int(n) + 2 * self.nbl for n in shape]
        if self.fs:                                  # model.py:102-109
            origin_pml[-1] = dtype(origin[-1])
            shape_pml[-1] -= self.nbl
        self.origin_pml, self.shape_pml = tuple(origin_pml), tuple(shape_pml)
        self.dim = len(shape)
        self._dt = dt
        if np.isscalar(vp):
            vp = np.full(shape, vp)
        self.vp = pad_edge(np.asarray(vp, dtype=dtype), self.nbl, self.fs)
        self.damp = init_damp(self.shape_pml, self.nbl, self.spacing, "damp", dtype, self.fs)

    @property
    def domain_size(self):
        return tuple((d - 1) * s for d, s in zip(self.shape, self.spacing))

    @property
    def critical_dt(self):
        return critical_dt(float(self.vp.max()), self.spacing, self.space_order, self.dtype,
                           self._dt)

    def update_vp(self, vp):
        vp = np.asarray(vp, dtype=self.dtype)
        self.vp = vp.copy() if vp.shape == self.shape_pml else pad_edge(vp, self.nbl, self.fs)

    def grid_struct(self, space_order=None):
        """(set ``self.kernel = 'OT4'`` for the fourth-order-in-time update; the caller passes dt = 1.73 x critical_dt)"""
        g = _Grid()
        g.ndim = self.dim
        g.space_order = int(space_order or self.space_order)
        for d in range(self.dim):
            g.shape[d] = self.shape_pml[d]
            g.spacing[d] = self.spacing[d]
            g.origin[d] = float(self.origin_pml[d])
        g.fs = int(self.fs)
        g.ot4 = int(getattr(self, 'kernel', 'OT2') == 'OT4')
        return g


# ----------------------------------------------------------------------------- sweeps
def _ptr(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct)) if a is not None else None


def _prep(model):
    dtype = np.dtype(model.dtype)
    ct = ctypes.c_float if dtype == np.float32 else ctypes.c_double
    sfx = "_f32" if dtype == np.float32 else "_f64"
    return dtype, ct, sfx


def forward(model, src_coords, rec_coords, src_data, nt, dt, save=False, u=None,
            space_order=None, time_m=1, time_M=None, vp=None, fast=False):
    """ForwardOperator.apply: returns (rec[nt, nrec], u) with u of shape (3|nt, *shape_pml)."""
    dtype, ct, sfx = _prep(model)
    g = model.grid_struct(space_order)
    vp = np.ascontiguousarray(model.vp if vp is None else vp, dtype=dtype)
    damp = np.ascontiguousarray(model.damp, dtype=dtype)
    src_coords = np.ascontiguousarray(np.reshape(src_coords, (-1, model.dim)), dtype=dtype)
    rec_coords = np.ascontiguousarray(np.reshape(rec_coords, (-1, model.dim)), dtype=dtype)
    nsrc, nrec = src_coords.shape[0], rec_coords.shape[0]
    src_data = np.ascontiguousarray(np.reshape(src_data, (nt, nsrc)), dtype=dtype)
    rec = np.zeros((nt, nrec), dtype=dtype)
    if u is None:
        u = np.zeros(((nt if save else 3),) + model.shape_pml, dtype=dtype)
    time_M = nt - 2 if time_M is None else time_M
    fn = getattr(lib(fast), "oracle_forward" + sfx)
    fn.restype = ctypes.c_int
    rc = fn(ctypes.byref(g), _ptr(vp, ct), _ptr(damp, ct), ctypes.c_double(float(dt)),
            int(nt), int(time_m), int(time_M), _ptr(src_data, ct), _ptr(src_coords, ct), nsrc,
            _ptr(rec, ct), _ptr(rec_coords, ct), nrec, _ptr(u, ct), int(bool(save)))
    if rc != 0:
        raise RuntimeError("oracle_forward failed: %d" % rc)
    return rec, u


def _adjoint_call(model, rec_data, rec_coords, nt, dt, u_hist, v, grad, srca, src_coords,
                  imaging, space_order, time_m, time_M, vp, fast):
    dtype, ct, sfx = _prep(model)
    g = model.grid_struct(space_order)
    vp = np.ascontiguousarray(model.vp if vp is None else vp, dtype=dtype)
    damp = np.ascontiguousarray(model.damp, dtype=dtype)
    rec_coords = np.ascontiguousarray(np.reshape(rec_coords, (-1, model.dim)), dtype=dtype)
    nrec = rec_coords.shape[0]
    rec_data = np.ascontiguousarray(np.reshape(rec_data, (nt, nrec)), dtype=dtype)
    nsrc = 0
    if src_coords is not None:
        src_coords = np.ascontiguousarray(np.reshape(src_coords, (-1, model.dim)), dtype=dtype)
        nsrc = src_coords.shape[0]
    if v is None:
        v = np.zeros((3,) + model.shape_pml, dtype=dtype)
    time_M = nt - 2 if time_M is None else time_M
    fn = getattr(lib(fast), "oracle_adjoint" + sfx)
    fn.restype = ctypes.c_int
    rc = fn(ctypes.byref(g), _ptr(vp, ct), _ptr(damp, ct), ctypes.c_double(float(dt)),
            int(nt), int(time_m), int(time_M), _ptr(rec_data, ct), _ptr(rec_coords, ct), nrec,
            _ptr(srca, ct), _ptr(src_coords, ct), nsrc, _ptr(u_hist, ct), _ptr(v, ct),
            _ptr(grad, ct), int(imaging))
    if rc != 0:
        raise RuntimeError("oracle_adjoint failed: %d" % rc)
    return v


def gradient(model, rec_data, rec_coords, u_hist, nt, dt, grad=None, v=None, space_order=None,
             time_m=1, time_M=None, vp=None, fast=False):
    """GradientOperator.apply: returns grad (accumulated into when given)."""
    dtype = np.dtype(model.dtype)
    if grad is None:
        grad = np.zeros(model.shape_pml, dtype=dtype)
    assert grad.dtype == dtype and grad.flags.c_contiguous
    u_hist = np.ascontiguousarray(u_hist, dtype=dtype)
    _adjoint_call(model, rec_data, rec_coords, nt, dt, u_hist, v, grad, None, None, 1,
                  space_order, time_m, time_M, vp, fast)
    return grad


def adjoint(model, rec_data, rec_coords, src_coords, nt, dt, v=None, space_order=None,
            time_m=1, time_M=None, vp=None, fast=False):
    """AdjointOperator.apply: returns (srca[nt, nsrc], v)."""
    dtype = np.dtype(model.dtype)
    nsrc = np.reshape(src_coords, (-1, model.dim)).shape[0]
    srca = np.zeros((nt, nsrc), dtype=dtype)
    v = _adjoint_call(model, rec_data, rec_coords, nt, dt, None, v, None, srca, src_coords, 0,
                      space_order, time_m, time_M, vp, fast)
    return srca, v


# ----------------------------------------------------------------------------- fwi.py glue
def least_square(x, y):
    """misfit/misfit.py:5-9."""
    r = x - y
    return .5 * np.linalg.norm(r.flatten()) ** 2, r


def fix_source_illumination(model, src_pos, rec_positions, g):
    """fwi.py:104-129, axis swap included (np.meshgrid(z, x))."""
    dx, dz = model.spacing
    nx, nz = model.shape
    if g.shape != (nx, nz):
        raise ValueError("Shape does not match!")
    x = np.arange(0, nx) * dx
    z = np.arange(0, nz) * dz
    xx, zz = np.meshgrid(z, x)
    sigma = dx + dz
    sx, sz = src_pos[0], src_pos[1]
    g = g * (1. - np.exp(-.5 * ((xx - sx) ** 2 + (zz - sz) ** 2) / (sigma ** 2)))
    for i in range(rec_positions.shape[0]):
        rx, rz = rec_positions[i, 0], rec_positions[i, 1]
        g = g * (1. - np.exp(-.5 * ((xx - rx) ** 2 + (zz - rz) ** 2) / (sigma ** 2)))
    return g


def fwi_obj_single(model, src_pos, rec_positions, wavelet, nt, dt, obs, misfit_func=least_square,
                   direct_wave=None, calc_grad=False, fast=False):
    """fwi.py:131-173 with resample == identity (the drivers' default)."""
    syn, wfd = forward(model, src_pos, rec_positions, wavelet, nt, dt, save=calc_grad, fast=fast)
    syn_data, obs_data = syn, obs
    if direct_wave is not None:
        syn_data = syn_data - direct_wave
        obs_data = obs_data - direct_wave
    fval, residual = misfit_func(syn_data, obs_data)
    residual = np.asarray(residual, dtype=model.dtype)
    illum, crop_grad = None, None
    if calc_grad:
        grad = gradient(model, residual, rec_positions, wfd, nt, dt, fast=fast)
        nbl = model.nbl
        sl = (slice(nbl, -nbl),) * model.dim
        crop_grad = fix_source_illumination(model, src_pos, rec_positions, np.array(grad)[sl])
        illum = (wfd * wfd).sum(axis=0)[sl]
        illum = fix_source_illumination(model, src_pos, rec_positions, illum)
    return fval, crop_grad, residual, illum


def fwi_obj_multi(model, src_positions, rec_positions, wavelet, nt, dt, obs, misfit_func=least_square,
                  direct_wave=None, mask=None, precond=True, calc_grad=False, fast=False):
    """fwi.py:175-205."""
    fval = .0
    grad = np.zeros(model.shape)
    illum = np.zeros(model.shape)
    residuals = []
    for i in range(src_positions.shape[0]):
        dw = direct_wave[i] if direct_wave is not None else None
        f_, g_, r_, il_ = fwi_obj_single(model, src_positions[i], rec_positions, wavelet, nt, dt,
                                         obs[i], misfit_func, dw, calc_grad, fast)
        fval += f_
        residuals.append(r_)
        if calc_grad:
            grad += g_
            illum += il_
    if calc_grad:
        if precond:
            grad /= np.sqrt(illum + 1e-30)
        if mask is not None:
            grad *= mask
    return fval, grad.reshape(-1).astype(np.float64), residuals
