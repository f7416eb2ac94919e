"""Few-shot / strong-scaling probe of the SM-resident engine: forward(+history, illumination) and adjoint+imaging
time of `nshots` concurrent shots for every feasible cluster size (1..16 SMs per shot)."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.filterwarnings("ignore")
import numpy as np
import torch
from devito_fwi_b200 import configs, resident, _lib
from devito_fwi_b200.wavesolver import grid_struct

torch.zeros(1, device='cuda')
which = sys.argv[1] if len(sys.argv) > 1 else 'marmousi'
shot_counts = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 4, 8, 29]
only_c = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else None
geom = {'marmousi': lambda: configs.marmousi()[1], 'marmousi2': lambda: configs.marmousi2()[1],
        'circle': lambda: configs.circle()[1]}[which]()
m = geom.model
g = grid_struct(m.grid, m.space_order)
ref_rec = {}
rows = []
for ns in shot_counts:
    shots = list(np.linspace(0, geom.nsrc - 1, ns).round().astype(int)) if ns < geom.nsrc else list(range(geom.nsrc))
    for C, P in [(c, pp) for c in range(1, resident.MAX_CLUSTER + 1) for pp in resident.ROWS_PER_THREAD]:
        if only_c and C not in only_c:
            continue
        p = resident.plan_exact(m.grid, m.space_order, m.nbl, C, P)
        if p is None:
            continue
        n = ctypes.c_int32()
        rc = _lib.lib().b2fwi_res2d_max_active_clusters(ctypes.byref(g), ctypes.byref(p), ctypes.byref(n))
        if rc != 0 or n.value <= 0:
            print(which, "ns=%d C=%d: occupancy query rc=%d n=%d" % (ns, C, rc, n.value), flush=True)
            continue
        if ns > 2 * n.value:
            continue
        try:
            sv = resident.ResidentSurvey(geom, shots, plan=p)
            rec = sv.forward(save=True, illum=True)
            res = rec.clone()
            sv.gradient(res)
            torch.cuda.synchronize()
            t = []
            for fn in (lambda: sv.forward(save=True, illum=True), lambda: sv.gradient(res)):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                t.append(e0.elapsed_time(e1) / 3)
            key = ns
            r0 = rec[0].cpu().numpy()
            gr = sv.grad[0].cpu().numpy().copy()
            if key not in ref_rec:
                ref_rec[key] = (r0, gr)
            same = bool(np.array_equal(ref_rec[key][0], r0) and np.array_equal(ref_rec[key][1], gr))
            print(which, "ns=%2d C=%2d P=%2d G=%2d T=%3d rows=%3d smem=%6d slots=%2d | fwd %.3f ms adj %.3f ms sum %.3f | per-shot %.3f ms | model %.2f | same=%s"
                  % (ns, C, p.rows_per_thread, p.groups, p.threads, p.rows_cta, p.smem_bytes, n.value, t[0], t[1], t[0] + t[1],
                     (t[0] + t[1]) / ns, resident.step_cost(p) * -(-ns // n.value), same), flush=True)
            rows.append(dict(cfg=which, ns=ns, C=C, P=p.rows_per_thread, T=p.threads, rows=p.rows_cta, slots=n.value,
                             fwd=t[0], adj=t[1]))
            del sv, rec, res
            torch.cuda.empty_cache()
        except Exception as e:
            print(which, "ns=%d C=%d failed: %r" % (ns, C, e), flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(rows, open('gpurun_out/probe_strong_%s.json' % which, 'w'))
