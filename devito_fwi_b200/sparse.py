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


ROW_TILE = 16            # rows per tile of the TMA sweep kernels (csrc/stream_tma.cu)
MAX_ROW_CON = 8          # contributions per (plane, row) up to which a sweep kernel stages the injection itself
MAX_PER_TILE = 64        # recorded points per plane and row tile up to which its service warps interpolate them


def fuse_tables(grid, corner_off, cell_off, cell_ptr):
    """Tables for the sweep kernels' service warps (b2fwi_sparse.con_* / pt_*): contributions bucketed by
    (plane, row), points by (plane, tile of ROW_TILE rows). Returns ({}, 0) for a map too dense for them (it then
    runs as separate kernels)."""
    npl, nr = grid.shape[0], grid.shape[1]
    sr = grid.slice_shape[2]
    sp = grid.slice_shape[1] * sr
    nrt = (nr + ROW_TILE - 1) // ROW_TILE
    valid = corner_off >= 0
    if not valid.any(axis=1).all():
        return {}, 0                 # a point with no corner inside the grid records 0: left to the stand-alone kernel
    con_off = np.repeat(np.asarray(cell_off, dtype=np.int64), np.diff(cell_ptr))
    row_bounds = (np.arange(npl, dtype=np.int64)[:, None] * sp + np.arange(nr, dtype=np.int64)[None, :] * sr).ravel()
    row_bounds = np.append(row_bounds, np.int64(npl) * sp)
    con_rowptr = np.searchsorted(con_off, row_bounds).astype(np.int32)
    max_row_con = int(np.diff(con_rowptr).max(initial=0))
    tile_bounds = (np.arange(npl, dtype=np.int64)[:, None] * sp +
                   np.arange(nrt, dtype=np.int64)[None, :] * (ROW_TILE * sr)).ravel()
    tile_bounds = np.append(tile_bounds, np.int64(npl) * sp)
    first = np.argmax(valid, axis=1)
    home = corner_off[np.arange(corner_off.shape[0]), first].astype(np.int64)
    order = np.argsort(home, kind='stable').astype(np.int32)
    pt_home = home[order]
    pt_rowptr = np.searchsorted(pt_home, tile_bounds).astype(np.int32)
    if max_row_con > MAX_ROW_CON or np.diff(pt_rowptr).max(initial=0) > MAX_PER_TILE:
        return {}, 0
    return dict(con_rowptr=con_rowptr, con_off=con_off, pt_order=order, pt_home=pt_home, pt_rowptr=pt_rowptr), max_row_con


class SparseMap(object):
    """Device-resident ``b2fwi_sparse`` for one set of point coordinates on one grid."""

    def __init__(self, grid, coords):
        import torch
        off, w = resolve(grid, coords)
        self.npoint, self.ncorner = off.shape
        flat_off = off.ravel()
        flat_w = w.ravel()
        pt = np.repeat(np.arange(self.npoint, dtype=np.int32), self.ncorner)
        # contributions of exactly zero weight (points on grid nodes / lines / planes) add +0 and are dropped
        valid = (flat_off >= 0) & (flat_w != 0)
        v_off, v_w, v_pt = flat_off[valid], flat_w[valid], pt[valid]
        order = np.lexsort((v_pt, v_off))           # by cell, then ascending point (stable in corner order)
        v_off, v_w, v_pt = v_off[order], v_w[order], v_pt[order]
        cells, start = np.unique(v_off, return_index=True)
        self.ncell = int(cells.size)
        cell_ptr = np.concatenate([start, [v_off.size]]).astype(np.int32)
        self.host = dict(corner_off=off, corner_w=w, cell_off=cells.astype(np.int64), cell_ptr=cell_ptr,
                         contrib_pt=v_pt.astype(np.int32), contrib_w=v_w.astype(np.float32))
        row_tile, max_row_con = 0, 0
        if grid.dim == 3 and HALO == 0:
            tables, max_row_con = fuse_tables(grid, off, self.host['cell_off'], cell_ptr)
            self.host.update(tables)
            row_tile = ROW_TILE if tables else 0
        dev = 'cuda'
        self._t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in self.host.items()}
        s = _lib.Sparse()
        s.npoint, s.ncorner, s.ncell = self.npoint, self.ncorner, self.ncell
        s.row_tile = row_tile
        s.max_row_con = max_row_con if row_tile else 0
        offs = np.concatenate([self.host['cell_off'], self.host.get('pt_home', np.zeros(0, np.int64))])
        s.z_min, s.r_min, s.p_min, s.z_max, s.r_max, s.p_max = 0, 0, 0, -1, -1, -1
        if offs.size and grid.dim == 3:
            sr_ = grid.slice_shape[2]
            sp_ = grid.slice_shape[1] * sr_
            zs, rs, ps = offs % sr_, (offs % sp_) // sr_, offs // sp_
            s.z_min, s.z_max, s.r_min, s.r_max = int(zs.min()), int(zs.max()), int(rs.min()), int(rs.max())
            s.p_min, s.p_max = int(ps.min()), int(ps.max())
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
