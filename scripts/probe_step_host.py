"""Host-side profile of one fwi_loss step on the Marmousi survey (what is left besides the two resident launches)."""
import sys, os, cProfile, pstats, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.filterwarnings("ignore")
import numpy as np, torch
import devito_fwi_b200 as b
from devito_fwi_b200 import fwi, configs
g_true, g_init, g_const, mask = configs.marmousi()
obs = fwi.fm_multi(g_true); dw = fwi.fm_multi(g_const)
model = g_init.model
x0 = (1. / (model.vp.data[model.nbl:-model.nbl, model.nbl:-model.nbl].astype(np.float64) ** 2)).ravel()
for _ in range(3):
    fwi.fwi_loss(x0, g_init, obs, fwi.least_square, dw, mask, True, True)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t0 = time.perf_counter(); fwi.fwi_loss(x0, g_init, obs, fwi.least_square, dw, mask, True, True); ts.append(time.perf_counter() - t0)
print("wall ms per step", [round(t * 1e3, 2) for t in ts])
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    fwi.fwi_loss(x0, g_init, obs, fwi.least_square, dw, mask, True, True)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22); print(s.getvalue()[:4500])
