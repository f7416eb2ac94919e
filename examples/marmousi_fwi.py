#!/usr/bin/env python
"""Marmousi / Marmousi2 FWI on the B200 path, structured like the reference's marmousi_fwi.py:62-181 and
marmousi2_fwi.py (L2, 1-D Wasserstein or QW2D misfit).

Observed data from the true model, direct wave from the water model, initial model = smooth_20, bathymetry
mask, illumination preconditioning, box constraints; the objective is `fwi.fwi_loss` (fwi.py:236-246).
The outer optimiser: the reference's own `minimize` + `optimize.NLCG` when the reference checkout is on
PYTHONPATH behind `devito_fwi_b200/compat`, otherwise SciPy's L-BFGS-B (the alternative the
reference itself documents at marmousi_fwi.py:165-171).

    python examples/marmousi_fwi.py --maxiter 10 --nsrc 29
    python examples/marmousi_fwi.py --config marmousi2 --misfit qw2d --maxiter 5
    torchrun --nproc-per-node 8 examples/marmousi_fwi.py       # shots sharded over the GPUs
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from devito_fwi_b200 import configs, dist, fwi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nsrc", type=int, default=29)
    ap.add_argument("--maxiter", type=int, default=10)
    ap.add_argument("--precond", type=int, default=1)
    ap.add_argument("--bathy", type=int, default=1)
    ap.add_argument("--odir", default="./result")
    ap.add_argument("--config", default="marmousi", choices=["marmousi", "marmousi2"],
                    help="marmousi_fwi.py (SMARMN, 29 shots) or marmousi2_fwi.py (SMARM2, 31 shots)")
    ap.add_argument("--misfit", default="l2", choices=["l2", "w1d", "qw2d"],
                    help="least squares, trace-by-trace 1-D Wasserstein, or the QW2D back-and-forth Wasserstein misfit "
                         "of marmousi2_fwi.py:131-132 (all three evaluated on the device)")
    args = ap.parse_args()
    import warnings
    warnings.filterwarnings("ignore")
    dist.init_from_env()
    rank0 = dist.rank() == 0

    if args.config == "marmousi2":
        g_true, g_init, g_const, mask = configs.marmousi2(nsrc=args.nsrc if args.nsrc != 29 else 31)
    else:
        g_true, g_init, g_const, mask = configs.marmousi(nsrc=args.nsrc)
    from devito_fwi_b200.misfit import qWasserstein
    misfit_func = {"l2": fwi.least_square, "w1d": qWasserstein(gamma=1.01, method='1d'),
                   "qw2d": qWasserstein(gamma=1.01, method='2d', num_steps=15, step_scale=4.)}[args.misfit]
    if not args.bathy:
        mask = None
    obs = fwi.fm_multi(g_true)                      # marmousi_fwi.py:120
    direct_wave = fwi.fm_multi(g_const)             # marmousi_fwi.py:128
    vmin, vmax = (1.5, 5.0) if args.config == "marmousi2" else (1.5, 5.2)      # marmousi2_fwi.py:155-156 / marmousi_fwi.py:154-155
    bounds = [1.0 / vmax ** 2, 1.0 / vmin ** 2]
    shape = g_init.model.shape
    nbl = g_init.model.nbl
    m0 = 1. / (g_init.model.vp.data[nbl:-nbl, nbl:-nbl].reshape(-1).astype(np.float64)) ** 2
    vp_true = g_true.model.vp.data[nbl:-nbl, nbl:-nbl]

    history = []

    def fun(x):
        f, g, _ = fwi.fwi_loss(x, g_init, obs, misfit_func, direct_wave, mask, bool(args.precond), True)
        history.append(f)
        return f, g

    tic = time.time()
    try:
        import minimize as ref_minimize            # the reference's, if its checkout is importable
        from optimize import NLCG
        log = os.path.join(args.odir, "log")
        opt = NLCG(ls_method='Bracket', step_len_init=0.05, max_ls=10, log_path=log, verbose=0)
        m = ref_minimize.minimize(opt, maxIter=args.maxiter, ftol=1e-3, gtol=1e-8, log_path=log).run(
            m0, g_init, obs, misfit_func, direct_wave, mask, bool(args.precond), bounds)
        driver = "reference minimize.py + optimize.NLCG"
        f_end = fwi.fwi_loss(m, g_init, obs, misfit_func, direct_wave, mask, bool(args.precond), False)[0]
        f_start = None
    except ImportError:
        from scipy.optimize import minimize as sp_minimize
        # the preconditioned gradient has an arbitrary scale: normalise the first step like the reference's
        # line search does (step_len_init relative to |p|)
        f0, g0 = fun(m0)
        # L-BFGS-B's first trial is a unit-norm step along -grad: scale the variables so that this step changes
        # the slowness by at most 5 % (what the reference's step_len_init does), and normalise f by f(m0)
        ghat = g0 / np.linalg.norm(g0)
        scale = 0.05 * np.max(np.abs(m0)) / np.max(np.abs(ghat))

        def scaled(y):
            f, g = fun(m0 + scale * y)
            return f / f0, g * (scale / f0)
        res = sp_minimize(scaled, np.zeros_like(m0), jac=True, method='L-BFGS-B',
                          bounds=[((bounds[0] - a) / scale, (bounds[1] - a) / scale) for a in m0],
                          options={'maxiter': args.maxiter, 'maxcor': 10, 'maxls': 10, 'gtol': 0., 'ftol': 1e-12})
        m = m0 + scale * res.x
        driver = "scipy L-BFGS-B"
        f_start, f_end = f0, float(res.fun) * f0
    toc = time.time()
    if rank0:
        vp = 1.0 / np.sqrt(m.reshape(shape))
        vp0 = 1. / np.sqrt(m0.reshape(shape))

        def rel_err(v, region):
            return np.linalg.norm((v - vp_true)[region]) / np.linalg.norm(vp_true[region])
        nx, nz = shape
        regions = [("whole model", (slice(None), slice(None))),
                   ("well illuminated (central 80 % in x, upper 40 % in z)", (slice(nx // 10, nx - nx // 10), slice(7, (2 * nz) // 5))),
                   ("deep part (lower 35 % in z)", (slice(None), slice((65 * nz) // 100, None)))]
        os.makedirs(args.odir, exist_ok=True)
        vp.astype(np.float32).tofile(os.path.join(args.odir, "marmousi_result_misfit_0"))
        print("driver: %s | %d objective evaluations in %.2f s on %d GPU(s)" % (driver, len(history), toc - tic,
                                                                              dist.world_size()))
        if f_start is not None:
            print("objective %.4e -> %.4e" % (f_start, f_end))
        for name, region in regions:
            print("relative model error vs true vp, %s: %.4f -> %.4f" % (name, rel_err(vp0, region), rel_err(vp, region)))
    return history


if __name__ == "__main__":
    main()
