import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.filterwarnings("ignore")
import torch
from devito_fwi_b200 import configs
from devito_fwi_b200.resident import ResidentSurvey
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, g, shots in (("marmousi29", configs.marmousi()[1], list(range(29))), ("marmousi5", configs.marmousi(nsrc=5)[1], list(range(5))),
                       ("circle11", configs.circle()[1], list(range(11))), ("marmousi2_31", configs.marmousi2()[1], list(range(31)))):
    for mc in (1, None):
        if mc is None:
            from devito_fwi_b200 import resident
            p = resident.plan_model(g.model.grid, g.model.space_order, g.model.nbl, 1)
            sv = ResidentSurvey(g, shots, min_cluster=p.cluster if p.cluster > 1 else 2) if p.cluster > 1 else None
            if sv is None: continue
            tag = "smallest-fit"
        else:
            sv = ResidentSurvey(g, shots); tag = "adaptive"
        rec = sv.forward(save=True, illum=True).clone()
        print("%-14s %-12s cluster=%d  fwd %.3f ms  adj %.3f ms" % (name, tag, sv.plan.cluster, t(lambda: sv.forward(save=True, illum=True)), t(lambda: sv.gradient(rec))), flush=True)
        del sv, rec; torch.cuda.empty_cache()
