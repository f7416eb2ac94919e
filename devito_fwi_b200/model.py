"""Physical model of the acoustic hot path (API of the reference's seismic/model.py).

``SeismicModel`` (= ``Model``) owns the padded grid (``shape + 2*nbl``), the edge-replicated
velocity ``vp`` [km/s], the sponge profile ``damp`` and the CFL time step.  Host logic only; the
fields are ``grid.Function`` objects whose device copies feed the CUDA kernels.

Reference lines mirrored: seismic/model.py:13-51 (initialize_damp), :91-158 (GenericModel),
:167-178 (_gen_phys_param), :269-283 (SeismicModel.__init__), :338-370 (CFL), :372-400 (update, m).
Elastic / TTI / visco parameters, free surface and ``smooth``/``dm`` are outside the hot path.
"""
import warnings
from math import factorial

import numpy as np

from .grid import Grid, Function, Constant, mmax, mmin

__all__ = ['SeismicModel', 'Model', 'initialize_damp', 'initialize_function']

_UNSUPPORTED = ('vs', 'epsilon', 'delta', 'theta', 'phi', 'qp', 'qs', 'lam', 'mu')


def _as_pairs(nbl, ndim):
    if np.isscalar(nbl):
        return [(int(nbl), int(nbl))] * ndim
    return [tuple(int(i) for i in p) for p in nbl]


def initialize_function(function, data, nbl):
    """devito.builtins.initialize_function(mode='constant'): copy ``data`` into the interior and
    replicate the edge values into the ``nbl`` layers, dimension by dimension."""
    pads = _as_pairs(nbl, function.grid.dim)
    function.data[...] = np.pad(np.asarray(data, dtype=function.dtype), pads, mode='edge')


def initialize_damp(damp, padsizes, spacing, abc_type="damp", fs=False):
    """Sponge profile of seismic/model.py:13-51: per dimension and side, layer cell ``i`` counted
    from the outer edge gets ``c*(pos - sin(2 pi pos)/(2 pi))/h`` with ``pos = |(nbl - i + 1)/nbl|``,
    ``c = 1.5 ln(1000)/nbl``; contributions of different dimensions add up in the corners.
    ``"mask"`` starts from 1 and subtracts instead."""
    sign = -1.0 if abc_type == "mask" else 1.0
    field = np.full(damp.grid.shape, 1.0 if abc_type == "mask" else 0.0, dtype=np.float64)
    profiles = []
    for d, ((nbl, nbr), h) in enumerate(zip(padsizes, spacing)):
        n = damp.grid.shape[d]
        prof = np.zeros(n)
        for width, left in ((nbl, True), (nbr, False)):
            if width == 0 or (left and fs and d == damp.grid.dim - 1):
                continue
            i = np.arange(width)
            pos = np.abs((width - i + 1) / float(width))
            val = sign * (1.5 * np.log(1.0 / 0.001) / width) * (pos - np.sin(2*np.pi*pos)/(2*np.pi)) / h
            if left:
                prof[:width] += val
            else:
                prof[n - width:] += val[::-1]
        field += prof.reshape([n if k == d else 1 for k in range(damp.grid.dim)])
        profiles.append(prof)
    damp.data[...] = field.astype(damp.dtype)
    # the profile is separable (a sum of 1-D profiles); the SM-resident engine uses that
    damp._profiles = profiles if abc_type == "damp" else None


def _second_derivative_weights(half_width):
    """Central FD weights of d2/dx2 on offsets -half_width..half_width (unit spacing)."""
    R = half_width
    side = [2.0 * (-1) ** (k + 1) * factorial(R) ** 2 / (k * k * factorial(R - k) * factorial(R + k))
            for k in range(1, R + 1)]
    return np.array(side[::-1] + [-2.0 * sum(side)] + side)


class GenericModel(object):
    """Grid + absorbing layer shared by all model flavours (seismic/model.py:87-224)."""

    def __init__(self, origin, spacing, shape, space_order, nbl=20,
                 dtype=np.float32, subdomains=(), bcs="damp", grid=None, fs=False):
        self.shape = tuple(shape)
        self.space_order = space_order
        self.nbl = int(nbl)
        self.origin = tuple([dtype(o) for o in origin])
        self.fs = fs
        origin_pml = [dtype(o - s*nbl) for o, s in zip(origin, spacing)]
        shape_pml = np.array(shape) + 2 * self.nbl
        if fs:
            # free surface: no absorbing layer above index 0 of the last dimension (seismic/model.py:102-109); the
            # top rows then use the mirrored stencil of operators.py:8-35 (grid.fs travels to the kernels)
            origin_pml[-1] = dtype(origin[-1])
            shape_pml[-1] -= self.nbl
        if grid is None:
            # physical extent is counted per cell, hence shape - 1 (seismic/model.py:113-117)
            extent = tuple(np.array(spacing) * (shape_pml - 1))
            self.grid = Grid(extent=extent, shape=shape_pml, origin=origin_pml, dtype=dtype,
                             subdomains=subdomains)
        else:
            self.grid = grid
        self.grid.fs = bool(fs)
        self._spacing = tuple(dtype(s) for s in spacing)
        self._physical_parameters = set()
        self.damp = None
        self._initialize_bcs(bcs=bcs)

    def _initialize_bcs(self, bcs="damp"):
        """(Re)build the damping field; flips an existing "mask" profile to "damp" (and back)
        with the reference's warning (seismic/model.py:126-149, wavesolver.py:30)."""
        if self.nbl == 0:
            self.damp = 1 if bcs == "mask" else 0
            return
        init = self.damp is None
        if init:
            self.damp = Function(name="damp", grid=self.grid)
            self._damp_type = None
        if callable(bcs):
            bcs(self.damp, self.nbl)
            self._damp_type = 'custom'
        else:
            re_init = ((bcs == "mask" and self._damp_type == "damp") or
                       (bcs == "damp" and self._damp_type == "mask"))
            if init or re_init:
                if re_init and not init:
                    bcs_o = "damp" if bcs == "mask" else "mask"
                    warnings.warn("Re-initializing damp profile from %s to %s" % (bcs_o, bcs))
                    warnings.warn("Model has to be created with `bcs=\"%s\"`"
                                  "for this WaveSolver" % bcs)
                initialize_damp(self.damp, self.padsizes, self.spacing, abc_type=bcs, fs=self.fs)
                self._damp_type = bcs
        self._physical_parameters.update(['damp'])

    @property
    def padsizes(self):
        pads = [(self.nbl, self.nbl) for _ in range(self.dim - 1)]
        pads.append((0 if self.fs else self.nbl, self.nbl))
        return pads

    def physical_params(self, **kwargs):
        known = [getattr(self, i) for i in self.physical_parameters]
        return {i.name: kwargs.get(i.name, i) or i for i in known}

    def _gen_phys_param(self, field, name, space_order, is_param=True, default_value=0):
        if field is None:
            return default_value
        if isinstance(field, np.ndarray):
            function = Function(name=name, grid=self.grid, space_order=space_order, parameter=is_param)
            initialize_function(function, field, self.padsizes)
        else:
            function = Constant(name=name, value=field, dtype=self.grid.dtype)
        self._physical_parameters.update([name])
        return function

    @property
    def physical_parameters(self):
        return tuple(self._physical_parameters)

    @property
    def dim(self):
        return self.grid.dim

    @property
    def spacing(self):
        return self.grid.spacing

    @property
    def space_dimensions(self):
        return tuple("xyz"[:self.dim]) if self.dim == 3 else ("x", "z")[:self.dim]

    @property
    def spacing_map(self):
        return self.grid.spacing_map

    @property
    def dtype(self):
        return self.grid.dtype

    @property
    def domain_size(self):
        """Physical size of the un-padded domain."""
        return tuple((d-1) * s for d, s in zip(self.shape, self.spacing))


class SeismicModel(GenericModel):
    """Acoustic model: ``vp`` in km/s, squared slowness ``m = 1/vp^2`` (seismic/model.py:227-400)."""

    def __init__(self, origin, spacing, shape, space_order, vp, nbl=20, fs=False,
                 dtype=np.float32, subdomains=(), bcs="mask", grid=None, **kwargs):
        for k in _UNSUPPORTED:
            if kwargs.get(k) is not None:
                raise NotImplementedError("parameter `%s`: only the isotropic acoustic model is "
                                          "implemented on the B200 path" % k)
        super(SeismicModel, self).__init__(origin, spacing, shape, space_order, nbl,
                                           dtype, subdomains, grid=grid, bcs=bcs, fs=fs)
        self.vp = self._gen_phys_param(vp, 'vp', space_order)
        self._vp_version = 0
        self._dt = kwargs.get('dt')       # user-prescribed time step (e.g. marmousi_fwi.py:68)
        self._dt_scale = 1

    @property
    def _max_vp(self):
        return mmax(self.vp)

    @property
    def _thomsen_scale(self):
        return 1

    @property
    def dt_scale(self):
        return self._dt_scale

    @dt_scale.setter
    def dt_scale(self, val):
        self._dt_scale = val

    @property
    def _cfl_coeff(self):
        """Courant number sqrt(a1 / (dim * sum|w|)), a1 = 4, with the weights of a 2*space_order-wide
        second-derivative stencil exactly as the reference takes them (seismic/model.py:350-353)."""
        w = _second_derivative_weights(self.space_order)
        return np.sqrt(4 / float(self.grid.dim * np.sum(np.abs(w))))

    @property
    def critical_dt(self):
        """CFL time step rounded to 4 digits; a user ``dt`` wins when admissible
        (seismic/model.py:355-370)."""
        dt = self._cfl_coeff * np.min(self.spacing) / (self._thomsen_scale*self._max_vp)
        dt = self.dtype("%.3e" % (self.dt_scale * dt))
        if self._dt:
            if self._dt > dt:
                raise ValueError("Critical dt: %f, set dt: %f" % (dt, self._dt))
            return self._dt
        return dt

    def update(self, name, value):
        """Replace a physical parameter; a ``self.shape`` array is re-padded (seismic/model.py:372-393)."""
        try:
            param = getattr(self, name)
        except AttributeError:
            setattr(self, name, self._gen_phys_param(value, name, self.space_order))
            return
        if isinstance(value, np.ndarray):
            if value.shape == param.shape:
                param.data[:] = value[:]
            elif value.shape == self.shape:
                initialize_function(param, value, self.nbl)
            else:
                raise ValueError("Incorrect input size %s for model" % (value.shape,) +
                                 " %s without or %s with padding" % (self.shape, param.shape))
        else:
            param.data = value
        self._vp_version += 1

    @property
    def m(self):
        """Squared slowness (values, not a symbol)."""
        vp = self.vp.data
        return 1 / (vp * vp)

    def smooth(self, physical_parameters, sigma=5.0):
        """Gaussian smoothing of parameters in place (devito.gaussian_smooth)."""
        from scipy.ndimage import gaussian_filter
        params = self.physical_params()
        for name in physical_parameters:
            f = params[name]
            f.data[...] = gaussian_filter(f.data, sigma=sigma, mode='nearest')


Model = SeismicModel
