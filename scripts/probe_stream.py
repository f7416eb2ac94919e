"""Quick timing probe of the streaming engine (not a benchmark): 3-D step kernels at 592^3 and a
Marmousi single-shot forward + gradient."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import devito_fwi_b200 as b
from devito_fwi_b200 import configs

def t3d(n, so, steps=12):
    geom = configs.layered3d(n=n, space_order=so, rec_decimate=4)
    model = geom.model
    solver = b.AcousticWaveSolver(model, geom, space_order=so)
    nt = geom.nt
    u = b.TimeFunction(name='u', grid=model.grid, time_order=2, space_order=so)
    solver.forward(u=u, time_M=4)
    rec, u, s = solver.forward(u=u, time_m=5, time_M=4 + steps)
    print("3D n=%d so=%d fwd: %s  frac-of-6456GB/s=%.3f" % (n, so, s, s.gbytess / 6455.9), flush=True)
    return s

for so in (8, 4, 16):
    t3d(512, so)
torch.cuda.empty_cache()
g_true, g_init, _, _ = configs.marmousi()
geom = b.fwi._shot_geometry(g_init, 14)
solver = b.AcousticWaveSolver(geom.model, geom, space_order=8)
for it in range(2):
    syn, u, s1 = solver.forward(save=True)
    res = b.Receiver(name='res', grid=geom.model.grid, time_range=geom.time_axis, coordinates=geom.rec_positions)
    res.data[:] = syn.data
    grad, s2 = solver.gradient(rec=res, u=u)
    print("marmousi shot:", s1, s2, flush=True)
