"""Sparse-point resolution: multilinear corner offsets / weights and the cell-centric CSR that the
injection kernels consume (struct b2fwi_sparse, include/b2fwi.h).

Follows the arithmetic of Devito's generated inject / interpolate sections
(reference: seismic/self_adjoint/sa_01_iso_implementation1.ipynb:1257-1315):
coordinates in grid dtype, ``idx = floor((coord - o)/h)``, ``p = (coord - o) - h*idx``, weights as the
expanded bilinear polynomials in 2-D, corner order x-major / z-minor.  Injection into the same cell
by several points is applied in ascending point order (deterministic, no atomics).
"""
import ctypes

import numpy as np

from . import _lib
from .grid import HALO

__all__ = ['SparseMap', 'sparse_map']


def resolve(grid, coords):
    """Host part: returns (corner_off[np, nc] int64, corner_w[np, nc] float32)."""
    f32 = np.float32
    coords = np.ascontiguousarray(np.reshape(coords, (-1, grid.dim)), dtype=f32)
    npoint, nd = coords.shape
    o = np.array(grid.origin, dtype=f32)
    h = np.array(grid.spacing, dtype=f32)
    rel = coords - o
    idx = np.floor(rel / h).astype(np.int64)
    p = rel - h * idx.astype(f32)
    # element strides of a pitched slice
    strides = [1]
    for n in grid.slice_shape[:0:-1]:
        strides.insert(0, strides[0] * n)
    strides = np.array(strides, dtype=np.int64)
    nc = 1 << nd
    off = np.empty((npoint, nc), dtype=np.int64)
    w = np.empty((npoint, nc), dtype=f32)
    if nd == 2:
        hx, hz = h
        px, pz = p[:, 0], p[:, 1]
        pp = px * pz / (hx * hz)
        w[:, 0] = pp - px / hx - pz / hz + f32(1)
        w[:, 1] = -pp + pz / hz
        w[:, 2] = -pp + px / hx
        w[:, 3] = pp
    else:
        fr = p / h
    c = 0
    shape = np.array(grid.shape)
    for corner in np.ndindex(*([2] * nd)):
        ci = idx + np.array(corner)
        valid = np.all((ci >= 0) & (ci < shape), axis=1)
        off[:, c] = np.where(valid, ((ci + HALO) * strides).sum(axis=1), -1)
        if nd == 3:
            wc = np.ones(npoint, dtype=f32)
            for d in range(3):
                wc = wc * (fr[:, d] if corner[d] else f32(1) - fr[:, d])
            w[:, c] = wc
        c += 1
    return off, w


class SparseMap(object):
    """Device-resident ``b2fwi_sparse`` for one set of point coordinates on one grid."""

    def __init__(self, grid, coords):
        import torch
        off, w = resolve(grid, coords)
        self.npoint, self.ncorner = off.shape
        flat_off = off.ravel()
        flat_w = w.ravel()
        pt = np.repeat(np.arange(self.npoint, dtype=np.int32), self.ncorner)
        valid = flat_off >= 0
        v_off, v_w, v_pt = flat_off[valid], flat_w[valid], pt[valid]
        order = np.lexsort((v_pt, v_off))           # by cell, then ascending point (stable in corner order)
        v_off, v_w, v_pt = v_off[order], v_w[order], v_pt[order]
        cells, start = np.unique(v_off, return_index=True)
        self.ncell = int(cells.size)
        cell_ptr = np.concatenate([start, [v_off.size]]).astype(np.int32)
        self.host = dict(corner_off=off, corner_w=w, cell_off=cells.astype(np.int64), cell_ptr=cell_ptr,
                         contrib_pt=v_pt.astype(np.int32), contrib_w=v_w.astype(np.float32))
        dev = 'cuda'
        self._t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in self.host.items()}
        s = _lib.Sparse()
        s.npoint, s.ncorner, s.ncell = self.npoint, self.ncorner, self.ncell
        for k, t in self._t.items():
            setattr(s, k, t.data_ptr() if t.numel() else None)
        self.struct = s

    def byref(self):
        return ctypes.byref(self.struct)


_CACHE = {}


def sparse_map(grid, coords):
    """Content-addressed cache: fwi.py rebuilds identical Receiver objects for every shot (fwi.py:153-155)."""
    coords = np.ascontiguousarray(np.reshape(coords, (-1, grid.dim)), dtype=np.float32)
    key = (grid._key(), coords.shape, coords.tobytes())
    m = _CACHE.get(key)
    if m is None:
        if len(_CACHE) > 256:
            _CACHE.clear()
        m = _CACHE[key] = SparseMap(grid, coords)
    return m
