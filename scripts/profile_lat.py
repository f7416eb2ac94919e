"""ncu target: `ns` Marmousi shots on clusters of `C` SMs (4-row-strip kernel when C >= 10): forward + gradient, twice."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.filterwarnings("ignore")
import numpy as np, torch
from devito_fwi_b200 import configs
from devito_fwi_b200.resident import ResidentSurvey
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 4
C = int(sys.argv[2]) if len(sys.argv) > 2 else 16
g = configs.marmousi()[1]
shots = list(np.linspace(0, g.nsrc - 1, ns).round().astype(int))
sv = ResidentSurvey(g, shots, min_cluster=C)
for _ in range(2):
    rec = sv.forward(save=True, illum=True)
    sv.gradient(rec.clone())
torch.cuda.synchronize()
print("done", sv.plan.cluster, sv.plan.rows_per_thread)
