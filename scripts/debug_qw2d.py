"""Stage-by-stage comparison of the device QW2D solver steps with the reference's fot2d.c functions: every step is fed
the REFERENCE's state, so the first diverging operation shows up without error amplification."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from devito_fwi_b200 import _lib
from oracle import ref_qw2d
from tests.golden.make_qw2d_golden import records
lib = _lib.lib(); R = ref_qw2d.lib()
nt, nrec = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (96, 33)
niter = int(sys.argv[3]) if len(sys.argv) > 3 else 4
n1, n2 = nrec, nt
f, g = records(nt, nrec, 3)
c = np.float32(-min(f.min(), g.min()) * np.float32(1.01))
mu = (f + c); mu = (mu / np.float32(mu.mean())).astype(np.float32)
nu = (g + c); nu = (nu / np.float32(nu.mean())).astype(np.float32)
y, x = np.meshgrid((np.arange(n2) + .5) / n2, (np.arange(n1) + .5) / n1, indexing='ij')
z = (0.5 * (x.astype(np.float32) ** 2 + y.astype(np.float32) ** 2)).astype(np.float32)
scratch = torch.empty(int(lib.b2fwi_qw2d_scratch_bytes(nt, nrec, 1)), dtype=torch.uint8, device='cuda')
P = lambda t: t.data_ptr() if t is not None else None
fp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
R.qw2d_ref_update.restype = ctypes.c_float
R.qw2d_ref_w2.restype = ctypes.c_float
def dev(op, a, b=None, c=None, d=None, sigma=0.0):
    A = torch.from_numpy(a.copy()).cuda(); B = torch.from_numpy(b.copy()).cuda() if b is not None else None
    C = torch.from_numpy(c.copy()).cuda() if c is not None else None; D = torch.from_numpy(d.copy()).cuda() if d is not None else None
    out = torch.zeros_like(A); scal = torch.zeros(4, device='cuda')
    _lib.check(lib.b2fwi_qw2d_debug_step(op, nt, nrec, P(A), P(B), P(C), P(D), ctypes.c_float(sigma), P(out), P(scal), P(scratch), None))
    torch.cuda.synchronize()
    return A.cpu().numpy(), out.cpu().numpy(), float(scal[0])
rel = lambda a, b: np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30)
def ref_dual(u):
    d = np.empty_like(u); R.qw2d_ref_dual(n1, n2, fp(np.ascontiguousarray(u)), fp(d)); return d
def ref_push(pot, dens):
    r = np.empty_like(pot); R.qw2d_ref_push(n1, n2, fp(pot), fp(dens), fp(r)); return r
sigma = np.float32(4.0) / max(mu.max(), nu.max())
phi, dual, rho = z.copy(), z.copy(), mu.copy()
old = R.qw2d_ref_w2(n1, n2, fp(phi), fp(dual), fp(mu), fp(nu))
def step_update(sigma, value, old, gsq):
    sigma, value, old, gsq = map(np.float32, (sigma, value, old, gsq))
    diff = value - old
    up = np.float32(1.0 / float(np.float32(.8)))
    if diff > gsq * sigma * np.float32(.75): return sigma * up
    if diff < gsq * sigma * np.float32(.25): return sigma * np.float32(.8)
    return sigma
for it in range(niter):
    for half, (pot_name, other, dens) in enumerate((("phi", nu, nu), ("dual", mu, mu))):
        pot = phi if half == 0 else dual
        pot_r = pot.copy(); h_r = R.qw2d_ref_update(n1, n2, fp(pot_r), fp(rho), fp(other), ctypes.c_float(sigma))
        pot_d, _, h_d = dev(1, pot, rho, other, sigma=float(sigma))
        print("it %d half %d update : delta rel %.2e  h1 %.7e vs %.7e" % (it, half, rel(pot_d - pot, pot_r - pot), h_d, h_r))
        a_r = ref_dual(pot_r); _, a_d, _ = dev(0, pot_r)
        b_r = ref_dual(a_r); _, b_d, _ = dev(0, a_r)
        print("            convexify: dual1 equal %s (max %.2e)  dual2 equal %s (max %.2e)" % (np.array_equal(a_d, a_r), np.abs(a_d - a_r).max(), np.array_equal(b_d, b_r), np.abs(b_d - b_r).max()))
        if half == 0: phi, dual = b_r, a_r
        else: dual, phi = b_r, a_r
        pot = phi if half == 0 else dual
        rho_r = ref_push(pot, dens); _, rho_d, _ = dev(2, pot, dens)
        print("            push     : rel %.2e max %.2e" % (rel(rho_d, rho_r), np.abs(rho_d - rho_r).max()))
        rho = rho_r
        v_r = R.qw2d_ref_w2(n1, n2, fp(phi), fp(dual), fp(mu), fp(nu)); _, _, v_d = dev(3, phi, dual, mu, nu)
        print("            w2       : %.9e vs %.9e" % (v_d, v_r))
        sigma = step_update(sigma, v_r, old, h_r); old = v_r
from scipy.fft import dctn
xr = np.random.default_rng(1).standard_normal((n2, n1)).astype(np.float32)
_, o4, _ = dev(4, xr)
_, o5, _ = dev(5, xr)
print("dct II  rel %.3e   dct III rel %.3e" % (rel(o4, dctn(xr.astype(np.float64), type=2)), rel(o5, dctn(xr.astype(np.float64), type=3))))

from scipy.fft import dct
_, o6, _ = dev(6, xr); _, o7, _ = dev(7, xr)
x64 = xr.astype(np.float64)
print("dct III along n1 (axis 1): rel %.3e ; along n2 (axis 0): rel %.3e" % (rel(o6, dct(x64, type=3, axis=1)), rel(o7, dct(x64, type=3, axis=0))))
r6 = dct(x64, type=3, axis=1)
print("row 0 dev", o6[0, :6], "ref", r6[0, :6])
print("row 5 dev", o6[5, :6], "ref", r6[5, :6])

_, w8, _ = dev(8, xr); _, W9, _ = dev(9, xr)
N = n1
def pre_np(X):
    w = np.empty(N)
    for k in range(N):
        kk = k if k <= N // 2 else N - k
        a = X[kk]; b = X[N - kk] if kk > 0 else 0.0
        cs, sn = np.cos(np.pi * kk / (2 * N)), np.sin(np.pi * kk / (2 * N))
        A = cs * a + sn * b; B = sn * a - cs * b
        if k > N // 2: B = -B
        w[k] = A - B
    return w
w_np = pre_np(x64[0])
print("pre  row0 max err %.3e" % np.abs(w8[0] - w_np).max())
Wnp = np.fft.rfft(w_np)
Wd = W9.ravel()[:2 * (N // 2 + 1)].reshape(-1, 2)
print("FFT  row0 max err re %.3e im %.3e" % (np.abs(Wd[:, 0] - Wnp.real).max(), np.abs(Wd[:, 1] - Wnp.imag).max()))
print("W dev[:4]", Wd[:4].tolist(), "np", [(c.real, c.imag) for c in Wnp[:4]])
