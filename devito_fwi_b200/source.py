"""Time axis and sparse point sources / receivers (API of the reference's seismic/source.py).

``PointSource`` (= ``Receiver`` = ``Shot``) keeps the SparseTimeFunction surface the drivers use:
``.data[nt, npoint]`` float32, ``.coordinates.data[npoint, dim]``, ``.time_values``, ``.resample``
(reference: seismic/source.py:15-75 TimeAxis, :78-177 PointSource, :181-351 wavelets).
"""
import numpy as np
from scipy import interpolate

__all__ = ['PointSource', 'Receiver', 'Shot', 'WaveletSource',
           'RickerSource', 'GaborSource', 'DGaussSource', 'TimeAxis']


class TimeAxis(object):
    """Uniform time axis defined by any three of start / step / num / stop.

    Same resolution rules as the reference (seismic/source.py:15-75): a missing ``num`` is
    ``ceil((stop - start + step)/step)`` and ``stop`` is then moved onto the last sample.
    """

    def __init__(self, start=None, step=None, num=None, stop=None):
        given = [v is not None for v in (start, step, num, stop)]
        if sum(given) != 3:
            raise ValueError("Three of args start, step, num and stop may be set")
        try:
            if start is None:
                start = stop - step*(num - 1)
            elif step is None:
                step = (stop - start)/(num - 1)
            elif num is None:
                num = int(np.ceil((stop - start + step)/step))
                stop = start + step*(num - 1)
            else:
                stop = start + step*(num - 1)
        except Exception:
            raise ValueError("Three of args start, step, num and stop may be set")
        if not isinstance(num, int):
            raise TypeError("input argument must be of type int")
        self.start, self.stop, self.step, self.num = start, stop, step, num
        self._tv = None

    def __str__(self):
        return "TimeAxis: start=%g, stop=%g, step=%g, num=%g" % \
               (self.start, self.stop, self.step, self.num)

    def _rebuild(self):
        return TimeAxis(start=self.start, stop=self.stop, num=self.num)

    @property
    def time_values(self):
        if self._tv is None:
            self._tv = np.linspace(self.start, self.stop, self.num)
        return self._tv


class _Coordinates(object):
    def __init__(self, data):
        self.data = data


class _SparseData(object):
    """[nt, npoint] float32 with a lazily synchronised device copy (see grid.DeviceBuffer)."""

    def __init__(self, nt, npoint, dtype):
        self.shape = (int(nt), int(npoint))
        self.dtype = np.dtype(dtype)
        self._host = None
        self._host_t = None
        self._dev = None
        self._newer = None
        self._hver = 0          # bumped whenever the host view is handed out (it may then be modified)

    def host(self):
        from .grid import cuda_available
        if self._host is None:
            if cuda_available() and self.dtype == np.float32:
                import torch
                self._host_t = torch.zeros(self.shape, dtype=torch.float32, pin_memory=True)
                self._host = self._host_t.numpy()
            else:
                self._host = np.zeros(self.shape, dtype=self.dtype)
            if self._dev is not None:
                self._newer = 'dev'
        if self._newer == 'dev':
            if self._host_t is not None:
                self._host_t.copy_(self._dev)
            else:
                self._host[...] = self._dev.cpu().numpy()
        self._newer = 'host'
        self._hver += 1
        return self._host

    def dev(self, write=False):
        import torch
        if self.dtype != np.float32:
            raise NotImplementedError("device compute is float32 only")
        if self._dev is None:
            self._dev = torch.zeros(self.shape, dtype=torch.float32, device='cuda')
            if self._host is not None:
                self._newer = 'host'
        if self._newer == 'host':
            self._dev.copy_(self._host_t if self._host_t is not None else torch.from_numpy(self._host))
            self._newer = None
        if write:
            self._newer = 'dev'
        return self._dev

    def adopt_dev(self, tensor):
        """Take ownership of a device tensor as the current value."""
        assert tuple(tensor.shape) == self.shape
        self._dev = tensor
        self._newer = 'dev'


class PointSource(object):
    """Set of sparse points with one time series each (seismic/source.py:78-174)."""

    def __init__(self, name=None, grid=None, time_range=None, npoint=None, data=None,
                 coordinates=None, coordinates_data=None, space_order=0, time_order=2, dtype=None,
                 **kwargs):
        if time_range is None:
            raise TypeError("Need `time_range`")
        if coordinates is None:
            coordinates = coordinates_data
        if npoint is None:
            if coordinates is None:
                raise TypeError("Need either `npoint` or `coordinates`")
            npoint = np.shape(coordinates)[0]
        self.name = name
        self.grid = grid
        self.npoint = int(npoint)
        self.nt = time_range.num
        self.dtype = dtype or (grid.dtype if grid is not None else np.float32)
        self._time_range = time_range._rebuild()
        ndim = grid.dim if grid is not None else np.shape(coordinates)[-1]
        coords = np.zeros((self.npoint, ndim), dtype=self.dtype)
        if coordinates is not None:
            coords[:] = np.reshape(coordinates, (self.npoint, ndim))
        self.coordinates = _Coordinates(coords)
        self._sdata = _SparseData(self.nt, self.npoint, self.dtype)
        if data is not None:
            self.data[:] = data

    @property
    def data(self):
        return self._sdata.host()

    @property
    def shape(self):
        return self._sdata.shape

    @property
    def time_values(self):
        return self._time_range.time_values

    @property
    def time_range(self):
        return self._time_range

    def resample(self, dt=None, num=None, rtol=1e-5, order=3):
        """Cubic-spline resampling onto a new step or sample count (seismic/source.py:140-170);
        returns ``self`` when the step does not change."""
        assert (dt is None) != (num is None), "exactly one of dt / num"
        axis = self._time_range
        if dt is None:
            target = TimeAxis(start=axis.start, stop=axis.stop, num=num)
            dt = target.step
        else:
            target = TimeAxis(start=axis.start, stop=axis.stop, step=dt)
        if np.isclose(dt, axis.step):
            return self
        old = self.data
        traces = np.zeros((target.num, old.shape[1]))
        for i in range(old.shape[1]):
            spline = interpolate.splrep(axis.time_values, old[:, i], k=order)
            traces[:, i] = interpolate.splev(target.time_values, spline)
        return PointSource(name=self.name, grid=self.grid, data=traces, time_range=target,
                           coordinates=self.coordinates.data)


Receiver = PointSource
Shot = PointSource


class WaveletSource(PointSource):
    """Sources carrying a pre-defined wavelet (seismic/source.py:181-245)."""

    def __init__(self, *args, **kwargs):
        kwargs.setdefault('npoint', 1)
        super(WaveletSource, self).__init__(*args, **kwargs)
        self.f0 = kwargs.get('f0')
        self.a = kwargs.get('a')
        self.t0 = kwargs.get('t0')
        for p in range(self.npoint):
            self.data[:, p] = self.wavelet

    @property
    def wavelet(self):
        raise NotImplementedError('Wavelet not defined')


class RickerSource(WaveletSource):
    """Ricker wavelet (seismic/source.py:248-277)."""

    @property
    def wavelet(self):
        t0 = self.t0 or 1 / self.f0
        a = self.a or 1
        r = (np.pi * self.f0 * (self.time_values - t0))
        return a * (1-2.*r**2)*np.exp(-r**2)


class GaborSource(WaveletSource):
    """Gabor wavelet (seismic/source.py:280-310)."""

    @property
    def wavelet(self):
        agauss = 0.5 * self.f0
        tcut = self.t0 or 1.5 / agauss
        s = (self.time_values - tcut) * agauss
        a = self.a or 1
        return a * np.exp(-2*s**2) * np.cos(2 * np.pi * s)


class DGaussSource(WaveletSource):
    """1st derivative of a Gaussian (seismic/source.py:313-351)."""

    @property
    def wavelet(self):
        t0 = self.t0 or 1 / self.f0
        a = self.a or 1
        time = (self.time_values - t0)
        return -2 * a * time * np.exp(- a * time**2)
