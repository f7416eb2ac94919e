"""Timing probe of the resident engine variants (not a benchmark)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.filterwarnings("ignore")
import torch
from devito_fwi_b200 import configs
from devito_fwi_b200.resident import ResidentSurvey
g = configs.marmousi()[1]
sv = ResidentSurvey(g, list(range(29)))
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("fwd save+illum  %.3f ms" % t(lambda: sv.forward(save=True, illum=True)))
print("fwd illum only  %.3f ms" % t(lambda: sv.forward(save=False, illum=True)))
print("fwd plain       %.3f ms" % t(lambda: sv.forward(save=False, illum=False)))
rec = sv.forward(save=True, illum=True).clone()
print("adj             %.3f ms" % t(lambda: sv.gradient(rec)))
