"""Devito-free data objects behind the reference's solver API.

Mirrors the small part of devito's type system that the hot path touches
(reference call sites: seismic/model.py:115-117,139,171-176; seismic/acoustic/wavesolver.py:104-106,
144-145,179-183; fwi.py:162): ``Grid``, ``Function``, ``TimeFunction``, ``Constant`` plus
``norm`` / ``mmax`` / ``mmin``.  Values live in a torch CUDA tensor laid out as the C ABI wants it
(pitched slices, include/b2fwi.h); ``.data`` is a host numpy view that is synchronised lazily.
Host-only use (no GPU) works for everything except launching kernels.
"""
import numpy as np

__all__ = ['Grid', 'Function', 'TimeFunction', 'Constant', 'norm', 'mmax', 'mmin', 'HALO']

HALO = 0          # device layout halo (see include/b2fwi.h); predicated loads synthesise the exterior
PITCH_ALIGN = 32  # floats


def _torch():
    import torch
    return torch


def cuda_available():
    try:
        return _torch().cuda.is_available()
    except Exception:   # pragma: no cover
        return False


class Grid(object):
    """Cartesian grid: ``shape`` points, physical ``extent``, ``origin`` (devito.Grid subset)."""

    def __init__(self, shape, extent=None, origin=None, dtype=np.float32, subdomains=()):
        self.shape = tuple(int(s) for s in shape)
        self.dim = len(self.shape)
        self.dtype = dtype
        extent = tuple(extent) if extent is not None else tuple(1.0 for _ in self.shape)
        self.extent = tuple(dtype(e) for e in extent)
        origin = tuple(origin) if origin is not None else tuple(0.0 for _ in self.shape)
        self.origin = tuple(dtype(o) for o in origin)
        self.spacing = tuple(dtype(e / (n - 1)) if n > 1 else dtype(e)
                             for e, n in zip(extent, self.shape))
        self.subdomains = {getattr(s, 'name', str(i)): s for i, s in enumerate(subdomains)}

    @property
    def spacing_map(self):
        return {}

    def _key(self):
        return (self.shape, self.extent, self.origin, np.dtype(self.dtype).str)

    def __eq__(self, other):
        return isinstance(other, Grid) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return "Grid[extent=%s, shape=%s, origin=%s]" % (self.extent, self.shape, self.origin)

    # ---- device layout (must agree with b2fwi_field_layout; checked in tests/test_capi.py)
    @property
    def pitch(self):
        n = self.shape[-1] + 2 * HALO
        return (n + PITCH_ALIGN - 1) // PITCH_ALIGN * PITCH_ALIGN

    @property
    def slice_shape(self):
        return tuple(n + 2 * HALO for n in self.shape[:-1]) + (self.pitch,)

    @property
    def slice_elems(self):
        return int(np.prod(self.slice_shape))


class DeviceBuffer(object):
    """``nlead`` leading entries (time slots or nothing) x one grid slice, host + device copies."""

    def __init__(self, grid, lead=(), dtype=np.float32):
        self.grid = grid
        self.lead = tuple(int(n) for n in lead)
        self.dtype = np.dtype(dtype)
        self._host = None
        self._dev = None
        self._newer = None     # None: in sync (or both unallocated == zero); 'host' | 'dev'

    @property
    def host_shape(self):
        return self.lead + self.grid.shape

    def _interior(self, t):
        idx = (Ellipsis,) + tuple(slice(HALO, HALO + n) for n in self.grid.shape)
        return t[idx]

    def _alloc_host(self):
        if cuda_available() and self.dtype == np.float32:
            torch = _torch()
            self._host_t = torch.zeros(self.host_shape, dtype=torch.float32, pin_memory=True)
            self._host = self._host_t.numpy()
        else:
            self._host = np.zeros(self.host_shape, dtype=self.dtype)
            self._host_t = None

    def host(self):
        """Host view; assumed to be modified by the caller."""
        if self._host is None:
            self._alloc_host()
            if self._dev is not None:
                self._newer = 'dev'
        if self._newer == 'dev':
            torch = _torch()
            src = self._interior(self._dev)
            if self._host_t is not None:
                self._host_t.copy_(src)
            else:
                self._host[...] = src.cpu().numpy()
            torch.cuda.current_stream().synchronize()
        self._newer = 'host'
        return self._host

    def host_ro(self):
        """Host view for reading only: brought up to date, but not marked as modified (so the next
        kernel call does not upload it again)."""
        newer = self._newer
        h = self.host()
        if newer != 'host':
            self._newer = None
        return h

    def reduce(self, op):
        """max / min over the domain; on the device when the device copy is current."""
        if self._dev is not None and self._newer != 'host':
            t = self._interior(self._dev)
            return float(t.max() if op == 'max' else t.min())
        h = self.host_ro()
        return float(np.max(h) if op == 'max' else np.min(h))

    def set_host(self, value):
        h = self.host()
        h[...] = value

    def dev(self, write=False):
        """Device tensor of shape lead + slice_shape (float32). ``write``: a kernel will modify it."""
        torch = _torch()
        if self.dtype != np.float32:
            raise NotImplementedError("device compute is float32 only (seismic/model.py:92 default dtype)")
        if not cuda_available():
            raise RuntimeError("devito_fwi_b200 needs a CUDA device: there is no CPU fallback")
        if self._dev is None:
            self._dev = torch.zeros(self.lead + self.grid.slice_shape, dtype=torch.float32, device='cuda')
            if self._host is not None:
                self._newer = 'host'
        if self._newer == 'host':
            src = self._host_t if self._host_t is not None else torch.from_numpy(self._host)
            self._interior(self._dev).copy_(src)
            self._newer = None
        if write:
            self._newer = 'dev'
        return self._dev

    def zero(self):
        if self._host is not None:
            self._host[...] = 0
        if self._dev is not None:
            self._dev.zero_()
        self._newer = None


class Function(object):
    """Time-invariant grid function (devito.Function subset)."""
    is_Constant = False

    def __init__(self, name=None, grid=None, space_order=1, parameter=False, dtype=None, **kwargs):
        self.name = name
        self.grid = grid
        self.space_order = space_order
        self.dtype = dtype or grid.dtype
        self._buf = DeviceBuffer(grid, (), self.dtype)

    @property
    def shape(self):
        return self.grid.shape

    @property
    def data(self):
        return self._buf.host()

    @data.setter
    def data(self, value):
        self._buf.set_host(value)

    def __repr__(self):
        return "%s(%s)" % (self.name, ", ".join("xyz"[:self.grid.dim] if self.grid.dim == 3 else "xz"))


class TimeFunction(object):
    """Time-varying grid function (devito.TimeFunction subset): ``save`` slots or a time_order+1 ring."""

    def __init__(self, name=None, grid=None, save=None, time_order=2, space_order=1, dtype=None, **kwargs):
        self.name = name
        self.grid = grid
        self.save = int(save) if save else None
        self.time_order = time_order
        self.space_order = space_order
        self.dtype = dtype or grid.dtype
        self.nslots = self.save if self.save else time_order + 1
        self._buf = DeviceBuffer(grid, (self.nslots,), self.dtype)

    @property
    def shape(self):
        return (self.nslots,) + self.grid.shape

    @property
    def data(self):
        return self._buf.host()

    @data.setter
    def data(self, value):
        self._buf.set_host(value)


class Constant(object):
    """Scalar parameter (devito.Constant subset)."""
    is_Constant = True

    def __init__(self, name=None, value=0., dtype=np.float32, **kwargs):
        self.name = name
        self.dtype = dtype
        self._value = dtype(value)

    @property
    def data(self):
        return self._value

    @data.setter
    def data(self, value):
        self._value = self.dtype(value)

    @property
    def value(self):
        return self._value


def _values(f):
    if isinstance(f, (Function, TimeFunction)):
        return f._buf.host_ro()
    if hasattr(f, 'data'):
        return np.asarray(f.data)
    return np.asarray(f)


def norm(f, order=2):
    """devito.norm: l2 norm of all values."""
    v = np.asarray(_values(f), dtype=np.float64).ravel()
    return f.dtype(np.linalg.norm(v, order)) if hasattr(f, 'dtype') and callable(f.dtype) \
        else np.linalg.norm(v, order)


def mmax(f):
    """devito.builtins.mmax"""
    if isinstance(f, (Function, TimeFunction)):
        return np.dtype(f.dtype).type(f._buf.reduce('max'))
    return np.max(_values(f))


def mmin(f):
    """devito.builtins.mmin"""
    if isinstance(f, (Function, TimeFunction)):
        return np.dtype(f.dtype).type(f._buf.reduce('min'))
    return np.min(_values(f))
