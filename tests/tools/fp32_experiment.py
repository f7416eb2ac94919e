"""Emulate the kernel's fp32 arithmetic variants in numpy and measure the distance to the fp64 oracle
(one Marmousi shot)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import ref
import devito_fwi_b200 as b
from devito_fwi_b200 import configs
from devito_fwi_b200.sparse import resolve
from tests.util import ref_model, rel_l2

f32 = np.float32
def fma(a, b_, c):   # a*b exact in fp64, one rounding of the sum (double rounding negligible)
    return (a.astype(np.float64) * np.float64(b_) + c.astype(np.float64)).astype(f32)

g_true, g_init, _, _ = configs.marmousi()
geom = b.fwi._shot_geometry(g_init, 14)
model = geom.model
model._initialize_bcs("damp")
rm = ref_model(model)
nt, dt = geom.nt, float(geom.dt)
wav = np.float64(geom.src.data)
s64, _ = ref.forward(rm, geom.src_positions, geom.rec_positions, wav, nt, dt)
rm32 = ref_model(model, np.float32)
s32, _ = ref.forward(rm32, geom.src_positions, geom.rec_positions, wav.astype(f32), nt, dt)
print("oracle fp32 vs fp64: %.3e" % rel_l2(s32, s64))

NX, NZ = model.grid.shape
R = 4
vp = model.vp.data.astype(np.float64); damp = model.damp.data.astype(np.float64)
m = 1/(vp*vp); den = m + dt*damp
h = 30.0
ck = ref.laplace_coeffs(8)     # double
off_s, w_s = resolve(model.grid, geom.src_positions)
off_r, w_r = resolve(model.grid, geom.rec_positions)
pitch = model.grid.pitch
def unoff(off): return off // pitch, off % pitch
sx, sz = unoff(off_s[0]); rx, rz = unoff(off_r)
dtf = f32(dt)

def run(variant, nsteps=None):
    c1 = (m/den).astype(f32)
    if variant in ('fold',):
        c2 = (dt*dt/den/(h*h)).astype(f32)
        wk = ck.astype(f32)                         # pure c_k
    else:
        c2 = (dt*dt/den).astype(f32)
        inv_h2 = f32(1)/(f32(h)*f32(h))
        wk = (ck.astype(f32) * inv_h2).astype(f32)
    if variant == 'current':
        c0 = f32(wk[0] + wk[0])
        c0_lo = f32(0)
    else:
        tot = -4.0*np.sum(wk[1:].astype(np.float64))
        c0 = f32(tot); c0_lo = f32(tot - np.float64(c0))
    wk_lo = np.zeros(R+1, f32)
    if variant == 'hilo':
        exact = ck/(h*h)
        wk = exact.astype(f32); wk_lo = (exact - wk.astype(np.float64)).astype(f32)
        tot = -4.0*np.sum(exact[1:]); c0 = f32(tot); c0_lo = f32(tot-np.float64(c0))
    P = R
    uc = np.zeros((NX+2*P, NZ+2*P), f32); up = np.zeros_like(uc)
    rec = np.zeros((nt, rx.shape[0]), f32)
    I = (slice(P, P+NX), slice(P, P+NZ))
    for t in range(1, (nsteps or nt-2)+1):
        C = uc[I]
        if variant == 'diff':
            lap = np.zeros_like(C)
            for k in range(1, R+1):
                for (a, b_) in ((uc[P+k:P+k+NX, P:P+NZ], uc[P-k:P-k+NX, P:P+NZ]), (uc[P:P+NX, P+k:P+k+NZ], uc[P:P+NX, P-k:P-k+NZ])):
                    lap = fma((a - C) + (b_ - C), wk[k], lap)
        else:
            lap = (c0 * C).astype(f32)
            if variant != 'current':
                lap = fma(C, c0_lo, lap)
            for k in range(1, R+1):   # rows (x) first then z, as the kernel
                a, b_ = uc[P+k:P+k+NX, P:P+NZ], uc[P-k:P-k+NX, P:P+NZ]
                lap = fma(a + b_, wk[k], lap)
                if variant == 'hilo': lap = fma(a + b_, wk_lo[k], lap)
            for k in range(1, R+1):
                a, b_ = uc[P:P+NX, P+k:P+k+NZ], uc[P:P+NX, P-k:P-k+NZ]
                lap = fma(a + b_, wk[k], lap)
                if variant == 'hilo': lap = fma(a + b_, wk_lo[k], lap)
        tt = fma(C - up[I], c1, C)
        un = fma(lap, c2, tt)
        # inject
        for c in range(4):
            v = f32(vp[sx[c], sz[c]])
            un[sx[c], sz[c]] += w_s[0, c] * f32(wav[t, 0]) * dtf * dtf * v * v
        # interp from uc
        acc = np.zeros(rx.shape[0], f32)
        for c in range(4):
            acc = acc + w_r[:, c] * C[rx[:, c], rz[:, c]]
        rec[t] = acc
        up[I] = un
        uc, up = up, uc
    return rec

for variant in ('current', 'c0fix', 'hilo', 'diff', 'fold'):
    r = run(variant)
    print("%-8s vs fp64: %.3e   vs oracle-fp32: %.3e" % (variant, rel_l2(r, s64), rel_l2(r, s32)), flush=True)
