"""BASELINE.json configs[4]: synthetic 3-D layered acoustic model 512^3 (+2*40 sponge = 592^3), so=8, one shot:
forward + checkpointed adjoint/imaging gradient on the streaming engine. Prints one JSON line."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.filterwarnings("ignore")
import numpy as np, torch
import devito_fwi_b200 as b
from devito_fwi_b200 import configs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
tn = float(sys.argv[2]) if len(sys.argv) > 2 else 1250.
dec = int(sys.argv[3]) if len(sys.argv) > 3 else 4
so = int(sys.argv[4]) if len(sys.argv) > 4 else 8
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
geom = configs.layered3d(n=n, space_order=so, tn=tn, rec_decimate=dec)
model = geom.model
npts = int(np.prod(model.grid.shape))
solver = b.AcousticWaveSolver(model, geom, space_order=so)
t0 = time.time()
res = b.Receiver(name='res', grid=model.grid, time_range=geom.time_axis, coordinates=geom.rec_positions)
for rep in range(2):   # shot 1 pays the one-off allocations (tens of GB of checkpoints), shot 2 is the steady state of a survey
    rec, cw, s_f = solver.forward(save='checkpoint')     # forward modelling + on-device checkpoints (pass 1)
    # residual = the data themselves (any record works for timing; parity is tested at small sizes)
    res._sdata.adopt_dev(rec._sdata.dev().clone())
    torch.cuda.synchronize()
    grad, s_g = solver.gradient(rec=res, u=cw)           # pass 2: recompute + adjoint/imaging (by parts)
    torch.cuda.synchronize()
    if rep == 0:
        first = round(s_f.time + s_g.time, 3)
        del cw, grad
steps = geom.nt - 2
gmax = float(grad._buf.dev().abs().max())
out = {"workload": "layered3d %d^3 (+2*%d) so=%d nt=%d, %d receivers" % (n, model.nbl, so, geom.nt, geom.nrec),
       "forward": {"s": round(s_f.time, 4), "gpts_per_s": round(s_f.gpointss, 1), "GBs_alg": round(s_f.gbytess, 1),
                   "frac_hbm": round(s_f.gbytess / peak, 3)},
       "gradient_checkpointed": {"s": round(s_g.time, 4),
                                 "sweeps": "recompute + adjoint/imaging (by parts)",
                                 "gpts_per_s_2sweeps": round(2 * npts * steps / s_g.time / 1e9, 1)},
       "shot_gradient_s": round(s_f.time + s_g.time, 3), "first_shot_gradient_s_incl_allocations": first,
       "shot_gradient": {"sweeps": "forward(+checkpoints) + recompute + adjoint/imaging (by parts)",
                         "GBs_alg_52B": round(52.0 * npts * steps / (s_f.time + s_g.time) / 1e9, 1),
                         "frac_hbm_52B": round(52.0 * npts * steps / (s_f.time + s_g.time) / 1e9 / peak, 3),
                         "GBs_moved_76B": round(76.0 * npts * steps / (s_f.time + s_g.time) / 1e9, 1)},
       "grad_absmax": gmax,
       "hbm_peak_alloc_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1), "wall_s": round(time.time() - t0, 1)}
print(json.dumps(out))
