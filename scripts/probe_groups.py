"""Marmousi2 (31 shots): objective+gradient time for different launch-group partitions."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.filterwarnings("ignore")
import numpy as np, torch
from devito_fwi_b200 import fwi, configs, resident
g_true, g_init, g_const, mask = configs.marmousi2()
obs = fwi.fm_multi(g_true); dw = fwi.fm_multi(g_const)
model = g_init.model; nbl = model.nbl
x0 = (1. / (model.vp.data[nbl:-nbl, nbl:-nbl].astype(np.float64) ** 2)).ravel()
auto = resident.partition_shots(model.grid, 8, nbl, 31)
out = {"auto": auto}
orig = resident.partition_shots
for name, groups in (("auto", None), ("31xC6", [(31, 6)]), ("22xC6+9xC8", [(22, 6), (9, 8)]), ("16xC8+15xC8", [(16, 8), (15, 8)]),
                     ("16xC7+15xC7", [(16, 7), (15, 7)]), ("11+10+10 xC8", [(11, 8), (10, 8), (10, 8)])):
    fwi._SURVEYS.clear()
    resident.partition_shots = (lambda *a, _g=groups, **k: _g) if groups else orig
    try:
        for _ in range(2):
            fwi.fwi_loss(x0, g_init, obs, fwi.least_square, dw, mask, True, True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            f, g, _ = fwi.fwi_loss(x0, g_init, obs, fwi.least_square, dw, mask, True, True)
        e1.record(); torch.cuda.synchronize()
        out[name] = {"ms": round(e0.elapsed_time(e1) / 3, 3), "f": f,
                     "groups": [[sv.nshots, int(sv.plan.cluster)] for sv in fwi._resident_surveys(g_init, list(range(31)))]}
    except Exception as e:
        out[name] = {"error": repr(e)[:120]}
resident.partition_shots = orig
print(json.dumps(out))
