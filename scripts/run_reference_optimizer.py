"""The reference's own optimiser driving the B200 objective: minimize.py (minimize.run) + optimize/NLCG (and LBFGS where
NumPy allows it) + misfit/least_square, all UNMODIFIED files of the reference checkout, on top of devito_fwi_b200/compat
(`import fwi` / `seismic` / `devito` resolve to this package), for a few iterations of circle_fwi.py:62-160.

    B2FWI_REFERENCE=/path/to/devito-fwi python scripts/run_reference_optimizer.py [iterations] [nsrc]

The reference checkout is not part of this repository (and does not exist on a fresh GPU box): the script exits with a
message when it is not found. Prints one line per iteration and a final JSON summary.
"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("B2FWI_REFERENCE", "/root/reference")


def main(iterations=2, nsrc=11, optimizer='NLCG'):
    if not os.path.isfile(os.path.join(REF, "minimize.py")):
        print("reference checkout not found at %s (set B2FWI_REFERENCE)" % REF)
        return None
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(ROOT, "devito_fwi_b200", "compat"))
    import warnings
    warnings.filterwarnings("ignore")
    import numpy as np
    import torch
    import fwi                                   # compat shim -> devito_fwi_b200.fwi
    import minimize as ref_minimize              # reference file: `from fwi import fwi_loss`
    import misfit as ref_misfit                  # reference package
    import optimize as ref_optimize              # reference package
    from devito_fwi_b200 import configs, _lib
    assert ref_minimize.__file__.startswith(REF) and ref_optimize.__file__.startswith(REF)
    assert ref_minimize.fwi_loss.__module__ == "devito_fwi_b200.fwi"
    lib = _lib.lib()

    geometry1, geometry0 = configs.circle(space_order=6, nsrc=nsrc)
    obs = fwi.fm_multi(geometry1, save=False)
    init_model = geometry0.model
    nbl = init_model.nbl
    v0 = init_model.vp.data[nbl:-nbl, nbl:-nbl]
    m0 = 1.0 / (v0.reshape(-1).astype(np.float64)) ** 2
    bounds = [1.0 / 4.0 ** 2, 1.0 / 2.5 ** 2]          # circle_fwi.py:137-139
    workdir = tempfile.mkdtemp(prefix="b2fwi_refopt_")
    os.chdir(workdir)
    log = os.path.join(workdir, "log")
    Opt = getattr(ref_optimize, optimizer)
    kw = dict(memory=10) if optimizer == 'LBFGS' else {}
    opt = Opt(ls_method='Bracket', step_len_init=0.05, max_ls=10, log_path=log, verbose=0, **kw)
    minimizer = ref_minimize.minimize(opt, maxIter=iterations, ftol=1e-9, gtol=1e-12, log_path=log)
    n0 = lib.b2fwi_launch_count()
    torch.cuda.synchronize()
    t0 = time.time()
    m = minimizer.run(m0.copy(), geometry0, obs, ref_misfit.least_square, None, None, True, bounds=bounds)
    torch.cuda.synchronize()
    wall = time.time() - t0
    # minimize.save_misfit appends "f |g|" per gradient evaluation
    hist = np.loadtxt(os.path.join(log, "misfit")).reshape(-1, 2)[:, 0] if os.path.exists(os.path.join(log, "misfit")) else []
    f_end, _, _ = fwi.fwi_loss(m, geometry0, obs, ref_misfit.least_square, None, None, True, calc_grad=False)
    v_true = geometry1.model.vp.data[nbl:-nbl, nbl:-nbl]
    err0 = float(np.linalg.norm(v0 - v_true))
    err1 = float(np.linalg.norm(1.0 / np.sqrt(m.reshape(v0.shape)) - v_true))
    out = {"optimizer": "reference optimize.%s + minimize.run (unmodified, %s)" % (optimizer, REF),
           "misfit": "reference misfit.least_square (recognised -> evaluated on the device)",
           "iterations": iterations, "shots": nsrc, "objective_history": [float(x) for x in np.atleast_1d(hist)],
           "objective_after": float(f_end), "model_error_before": err0, "model_error_after": err1,
           "wall_s": round(wall, 3), "gpu_launches": int(lib.b2fwi_launch_count() - n0),
           "numpy": np.__version__}
    print(json.dumps(out))
    return out


if __name__ == "__main__":
    it = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    ns = int(sys.argv[2]) if len(sys.argv) > 2 else 11
    for name in ("NLCG", "LBFGS"):
        try:
            main(it, ns, name)
        except Exception as e:      # the reference's LBFGS core compares a memmap with [] (NumPy >= 2 rejects it)
            print(json.dumps({"optimizer": name, "error": repr(e)[:300]}))
