"""Phase timing of the full-size 3-D shot (592^3, so=8, nt=690): ring-only forward, checkpointed forward (cold / warm
allocator), gradient pass 2; with SM clock samples."""
import sys, os, json, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.filterwarnings("ignore")
import numpy as np, torch
import devito_fwi_b200 as b
from devito_fwi_b200 import configs

tn = float(sys.argv[1]) if len(sys.argv) > 1 else 1250.
geom = configs.layered3d(n=512, space_order=8, tn=tn, rec_decimate=4)
model = geom.model
solver = b.AcousticWaveSolver(model, geom, space_order=8)
clk = []
stop = False
def sample():
    while not stop:
        try:
            o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                               stdout=subprocess.PIPE, text=True, timeout=5).stdout.strip().split(",")
            clk.append((time.time(), float(o[0]), float(o[1])))
        except Exception:
            pass
        time.sleep(0.05)
th = threading.Thread(target=sample, daemon=True); th.start()
out = {}
def phase(name, fn):
    torch.cuda.synchronize(); t0 = time.time(); n0 = len(clk)
    r = fn(); torch.cuda.synchronize(); t1 = time.time()
    c = [x[1] for x in clk if t0 <= x[0] <= t1]; p = [x[2] for x in clk if t0 <= x[0] <= t1]
    out[name] = {"s": round(t1 - t0, 4), "sm_mhz": (min(c), max(c)) if c else None, "power_w": max(p) if p else None}
    return r
phase("forward_ring", lambda: solver.forward())
phase("forward_ring_again", lambda: solver.forward())
rec, cw, _ = phase("forward_checkpoint_cold", lambda: solver.forward(save='checkpoint'))
del cw
rec, cw, _ = phase("forward_checkpoint_warm", lambda: solver.forward(save='checkpoint'))
res = b.Receiver(name='res', grid=model.grid, time_range=geom.time_axis, coordinates=geom.rec_positions)
res._sdata.adopt_dev(rec._sdata.dev().clone())
phase("gradient_pass2", lambda: solver.gradient(rec=res, u=cw))
phase("gradient_pass2_again", lambda: solver.gradient(rec=res, u=cw))
stop = True
out["steps"] = geom.nt - 2
out["kept_steps"], out["segment"], out["segments"], out["cw_GB"] = cw.nkeep_steps, cw.S, len(cw.segs), round(cw.nbytes / 1e9, 1)
print(json.dumps(out))
