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
    # 8 time steps in 2 checkpoint segments: forward (4 on the ring + 4 in history mode), then adjoint+imaging (4),
    # recompute (4), adjoint+imaging (4) -- every TMA kernel variant of the 3-D shot gradient
    rec, cw, _ = solver.forward(save='checkpoint', time_M=8, segment=4, keep_segments=1)
    res = b.Receiver(name='res', grid=geom.model.grid, time_range=geom.time_axis, coordinates=geom.rec_positions)
    res._sdata.adopt_dev(rec._sdata.dev().clone())
    solver.gradient(rec=res, u=cw, time_M=8)
    torch.cuda.synchronize()
print("done")
