import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.filterwarnings("ignore")
import torch
from devito_fwi_b200 import configs, resident, _lib
from devito_fwi_b200.wavesolver import grid_struct
torch.zeros(1, device='cuda')
for name, geom in (('marmousi', configs.marmousi()[1]), ('marmousi2', configs.marmousi2()[1]), ('circle', configs.circle()[1])):
    m = geom.model
    g = grid_struct(m.grid, m.space_order)
    for C in range(1, 9):
        p = resident.plan_model(m.grid, m.space_order, m.nbl, min_cluster=C)
        if p is None or p.cluster != C: continue
        n = ctypes.c_int32()
        rc = _lib.lib().b2fwi_res2d_max_active_clusters(ctypes.byref(g), ctypes.byref(p), ctypes.byref(n))
        print(name, "C=%d P=%d G=%d T=%d rows=%d smem=%d -> max clusters %d (rc %d)" % (C, p.rows_per_thread, p.groups, p.threads, p.rows_cta, p.smem_bytes, n.value, rc), flush=True)
