"""Small driver for ncu: one Marmousi 29-shot resident forward + backward, one 3-D streaming step pair."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.filterwarnings("ignore")
import numpy as np, torch
import devito_fwi_b200 as b
from devito_fwi_b200 import configs
from devito_fwi_b200.resident import ResidentSurvey
what = sys.argv[1] if len(sys.argv) > 1 else "2d"
if what == "2d":
    g = configs.marmousi()[1]
    sv = ResidentSurvey(g, list(range(29)))
    for _ in range(2):
        rec = sv.forward(save=True, illum=True)
        sv.gradient(rec.clone())
    torch.cuda.synchronize()
else:
    geom = configs.layered3d(n=512, space_order=8, rec_decimate=8)
    solver = b.AcousticWaveSolver(geom.model, geom, space_order=8)
    u = b.TimeFunction(name='u', grid=geom.model.grid, time_order=2, space_order=8)
    solver.forward(u=u, time_M=8)
    torch.cuda.synchronize()
print("done")
