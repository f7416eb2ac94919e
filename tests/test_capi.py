"""C-ABI checks that need no GPU: the library loads, exports every symbol include/b2fwi.h declares,
agrees with the Python-side layout rule and rejects bad arguments with the documented codes."""
import ctypes
import os
import re

import numpy as np
import pytest

import __graft_entry__ as entry
from devito_fwi_b200 import _lib
from devito_fwi_b200.grid import Grid, HALO
from devito_fwi_b200.wavesolver import grid_struct

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    entry.build()
    return _lib.lib()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "b2fwi.h")).read()
    declared = set(re.findall(r"\b(b2fwi_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), "missing export %s" % name
    assert declared == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with the header"
    assert lib.b2fwi_version() == 101


@pytest.mark.parametrize("shape", [(380, 186), (281, 281), (75, 67), (592, 592, 592), (34, 29, 38)])
def test_layout_matches_python(lib, shape):
    grid = Grid(shape=shape, extent=tuple(10. * (n - 1) for n in shape))
    g = grid_struct(grid, 8)
    stride = (ctypes.c_int64 * 3)()
    base, elems = ctypes.c_int64(), ctypes.c_int64()
    assert lib.b2fwi_field_layout(ctypes.byref(g), ctypes.byref(stride), ctypes.byref(base),
                                  ctypes.byref(elems)) == 0
    assert elems.value == grid.slice_elems
    assert grid.pitch % 32 == 0 and grid.pitch >= shape[-1]
    expect = [int(np.prod(grid.slice_shape[d + 1:])) for d in range(len(shape))]
    assert list(stride)[:len(shape)] == expect
    assert base.value == sum(HALO * s for s in expect)


def test_argument_errors(lib):
    grid = Grid(shape=(40, 40), extent=(390., 390.))
    g = grid_struct(grid, 8)
    g.ndim = 4
    assert lib.b2fwi_field_layout(ctypes.byref(g), None, None, None) == -1
    assert b"ndim" in lib.b2fwi_last_error()
    g = grid_struct(grid, 8)
    g.space_order = 7
    assert lib.b2fwi_field_layout(ctypes.byref(g), None, None, None) == -1
    g = grid_struct(grid, 18)
    assert lib.b2fwi_field_layout(ctypes.byref(g), None, None, None) == -1
    g = grid_struct(grid, 8)
    # time range outside [1, nt-2] is rejected before any launch
    rc = lib.b2fwi_forward(ctypes.byref(g), 1, 1, ctypes.c_float(1.0), 10, 0, 8, None, None, None, None,
                           1, 0, None, None, 0, None)
    assert rc == -1 and b"time range" in lib.b2fwi_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    rc = lib.b2fwi_forward(ctypes.byref(g), None, None, ctypes.c_float(1.0), 10, 1, 8, None, None, None,
                           None, None, 0, None, None, 0, None)
    assert rc == -1 and b"NULL" in lib.b2fwi_last_error()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import devito_fwi_b200 as b
    model = b.demo_model('constant-isotropic', shape=(21, 21), nbl=4, space_order=4)
    geom = b.setup_geometry(model, 50.)
    solver = b.AcousticWaveSolver(model, geom, space_order=4)
    with pytest.raises((RuntimeError, AssertionError)):
        solver.forward()
