import sys, os, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.filterwarnings("ignore")
import numpy as np, torch, time
import devito_fwi_b200 as b
from devito_fwi_b200 import configs
geom = configs.layered3d(n=512, space_order=8, tn=20., rec_decimate=4)
solver = b.AcousticWaveSolver(geom.model, geom, space_order=8)
solver.forward(); torch.cuda.synchronize()
for k in range(2):
    t0 = time.time(); solver.forward(); torch.cuda.synchronize(); print("forward wall", time.time() - t0, "steps", geom.nt - 2)
pr = cProfile.Profile(); pr.enable(); solver.forward(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(18); print(s.getvalue()[:3500])
