import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.filterwarnings("ignore")
import torch
from devito_fwi_b200 import configs, _lib
from devito_fwi_b200.resident import ResidentSurvey
from devito_fwi_b200.wavesolver import _ptr, _stream
g = configs.marmousi()[1]
sv = ResidentSurvey(g, list(range(29)))
sv.set_model()
def fwd(rec):
    _lib.check(_lib.lib().b2fwi_res2d_forward(ctypes.byref(sv.gs), ctypes.byref(sv.plan), _ptr(sv.B), _ptr(sv.sx), _ptr(sv.sz),
        ctypes.c_float(sv.dt), sv.nt, 1, sv.nt - 2, sv.nshots, _ptr(sv.src), 1, sv.maps_fwd.byref(),
        _ptr(sv.rec) if rec else None, sv.nrec, None, None, _stream()))
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("fwd plain with recording    %.3f ms" % t(lambda: fwd(True)))
print("fwd plain without recording %.3f ms" % t(lambda: fwd(False)))
