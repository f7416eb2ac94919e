"""Per-window timing of the 3-D forward sweep (592^3, so=8): ms/step as the wavefield fills the grid."""
import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.filterwarnings("ignore")
import numpy as np, torch
import devito_fwi_b200 as b
from devito_fwi_b200 import configs, _lib
from devito_fwi_b200.wavesolver import _ptr, _stream
from devito_fwi_b200.sparse import sparse_map

geom = configs.layered3d(n=512, space_order=8, tn=1250., rec_decimate=4)
model = geom.model
solver = b.AcousticWaveSolver(model, geom, space_order=8)
lib = _lib.lib(); g = solver._gs(); dt = float(solver.dt)
vp_dev = solver._vp_dev(model.vp); coef = solver._coeffs(vp_dev, dt)
src = geom.src; src_map = sparse_map(model.grid, src.coordinates.data); src_dev = src._sdata.dev()
ring = torch.zeros((3,) + model.grid.slice_shape, dtype=torch.float32, device='cuda')
nt = geom.nt
W = 43
res = []
for rep in range(2):
    ring.zero_()
    row = []
    for ta in range(1, nt - 1, W):
        tb = min(ta + W - 1, nt - 2)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.b2fwi_forward(ctypes.byref(g), _ptr(vp_dev), _ptr(coef), ctypes.c_float(dt), nt, ta, tb,
                                     _ptr(src_dev), src_map.byref(), None, None, _ptr(ring), 0, None, None, 0, _stream()))
        e1.record(); e1.synchronize()
        row.append(round(e0.elapsed_time(e1) / (tb - ta + 1), 4))
    res.append(row)
    nz = float((ring[0] != 0).float().mean())
print(json.dumps({"ms_per_step_by_window": res, "nonzero_fraction_end": nz}))
