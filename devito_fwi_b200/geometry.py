"""Acquisition geometry (API of the reference's seismic/utils.py:12-189)."""
import warnings

import numpy as np

from .source import (PointSource, Receiver, RickerSource, GaborSource, WaveletSource,
                     DGaussSource, TimeAxis)

__all__ = ['AcquisitionGeometry', 'setup_geometry', 'setup_rec_coords', 'sources']

sources = {'Wavelet': WaveletSource, 'Ricker': RickerSource, 'Gabor': GaborSource,
           'DGauss': DGaussSource}


def setup_rec_coords(model):
    """One receiver per grid column (2-D) / per (x, y) node (3-D) at depth 2*h (seismic/utils.py:27-47)."""
    nrecx = model.shape[0]
    recx = np.linspace(model.origin[0], model.domain_size[0], nrecx)
    if model.dim == 1:
        return recx.reshape((nrecx, 1))
    depth = model.origin[-1] + 2 * model.spacing[-1]
    if model.dim == 2:
        rec = np.empty((nrecx, 2))
        rec[:, 0] = recx
        rec[:, 1] = depth
        return rec
    nrecy = model.shape[1]
    recy = np.linspace(model.origin[1], model.domain_size[1], nrecy)
    rec = np.empty((nrecx * nrecy, 3))
    rec[:, 0] = np.repeat(recx, nrecy)
    rec[:, 1] = np.tile(recy, nrecx)
    rec[:, 2] = depth
    return rec


def setup_geometry(model, tn, f0=0.010):
    """Source in the middle of the domain one cell below the surface (seismic/utils.py:12-24)."""
    src = np.empty((1, model.dim))
    src[0, :] = np.array(model.domain_size) * .5
    if model.dim > 1:
        src[0, -1] = model.origin[-1] + model.spacing[-1]
    return AcquisitionGeometry(model, setup_rec_coords(model), src, t0=0.0, tn=tn,
                               src_type='Ricker', f0=f0)


class AcquisitionGeometry(object):
    """Source / receiver positions plus the time axis of one survey (seismic/utils.py:50-189)."""

    def __init__(self, model, rec_positions, src_positions, t0, tn, **kwargs):
        self.rec_positions = np.reshape(rec_positions, (-1, model.dim))
        self.src_positions = np.reshape(src_positions, (-1, model.dim))
        self._nrec = self.rec_positions.shape[0]
        self._nsrc = self.src_positions.shape[0]
        self._src_type = kwargs.get('src_type')
        assert (self.src_type in sources or self.src_type is None)
        self._f0 = kwargs.get('f0')
        self._a = kwargs.get('a', None)
        self._t0w = kwargs.get('t0w', None)
        if self._src_type is not None and self._f0 is None:
            raise ValueError("Peak frequency must be provided in KH for source of type %s"
                             % self._src_type)
        self._grid = model.grid
        self._model = model
        self._dt = model.critical_dt
        self._t0 = t0
        self._tn = tn
        self._src_data = kwargs.get('src_data', None)
        self._filter = kwargs.get('filter', None)

    def resample(self, dt):
        self._dt = dt
        return self

    @property
    def time_axis(self):
        return TimeAxis(start=self.t0, stop=self.tn, step=self.dt)

    src_type = property(lambda self: self._src_type)
    grid = property(lambda self: self._grid)
    model = property(lambda self: self._model)
    f0 = property(lambda self: self._f0)
    tn = property(lambda self: self._tn)
    t0 = property(lambda self: self._t0)
    dt = property(lambda self: self._dt)
    nrec = property(lambda self: self._nrec)
    nsrc = property(lambda self: self._nsrc)

    @property
    def nt(self):
        return self.time_axis.num

    @property
    def dtype(self):
        return self.grid.dtype

    @property
    def rec(self):
        return self.new_rec()

    def new_rec(self, name='rec'):
        return Receiver(name=name, grid=self.grid, time_range=self.time_axis, npoint=self.nrec,
                        coordinates=self.rec_positions)

    @property
    def adj_src(self):
        if self.src_type is None:
            warnings.warn("No source type defined, returning uninitiallized (zero) shot record")
            return self.new_rec()
        adj_src = sources[self.src_type](name='rec', grid=self.grid, f0=self.f0,
                                         time_range=self.time_axis, npoint=self.nrec,
                                         coordinates=self.rec_positions, t0=self._t0w, a=self._a)
        # time-reversed wavelet: a proper shot record instead of zeros
        for i in range(self.nrec):
            adj_src.data[:, i] = adj_src.wavelet[::-1]
        return adj_src

    @property
    def src(self):
        return self.new_src()

    def new_src(self, name='src', src_type='self'):
        if self.src_type is None or src_type is None:
            return PointSource(name=name, grid=self.grid, time_range=self.time_axis,
                               npoint=self.nsrc, coordinates=self.src_positions)
        source = sources[self.src_type](name=name, grid=self.grid, f0=self.f0,
                                        time_range=self.time_axis, npoint=self.nsrc,
                                        coordinates=self.src_positions, t0=self._t0w, a=self._a)
        if self._filter is not None:
            self._filter.df = 1000 / self._dt
            for i in range(self.nsrc):
                source.data[:, i] = self._filter(source.data[:, i])
        return source
