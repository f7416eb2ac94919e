"""The reference's own optimiser / misfit packages, unmodified, on top of the compat shims
(INTEGRATION.md). Needs the reference checkout (this container only; skipped on the GPU box)."""
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "devito_fwi_b200", "compat")

needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


@pytest.fixture()
def ref_path(monkeypatch, tmp_path):
    monkeypatch.syspath_prepend(REF)
    monkeypatch.syspath_prepend(COMPAT)
    monkeypatch.chdir(tmp_path)
    for m in ("fwi", "seismic", "devito", "w2", "minimize", "optimize", "misfit"):
        sys.modules.pop(m, None)
    yield
    for m in list(sys.modules):
        if m.split(".")[0] in ("fwi", "seismic", "devito", "w2", "minimize", "optimize", "misfit", "examples"):
            sys.modules.pop(m, None)


@needs_ref
def test_reference_modules_import_on_the_shims(ref_path):
    import fwi
    import minimize          # reference file, does `from fwi import fwi_loss`
    import misfit            # reference package, needs the w2 shim
    import optimize          # reference package
    from seismic import Model, Receiver, AcquisitionGeometry
    from seismic.acoustic import AcousticWaveSolver
    from devito import Function
    assert minimize.fwi_loss is fwi.fwi_loss
    assert fwi.fwi_loss.__module__ == "devito_fwi_b200.fwi"
    assert AcousticWaveSolver.__module__ == "devito_fwi_b200.wavesolver"
    assert Function.__module__ == "devito_fwi_b200.grid"
    f, r = misfit.least_square(np.ones((4, 3), np.float32), np.zeros((4, 3), np.float32))
    assert np.isclose(f, 6.0) and r.shape == (4, 3)
    import devito_fwi_b200.fwi as ours
    assert ours._is_l2(misfit.least_square)          # recognised -> on-device misfit


@needs_ref
def test_reference_minimize_runs_unchanged_on_the_fg_contract(ref_path, tmp_path):
    """minimize.py + optimize/NLCG + bracketing line search, unmodified, driven through the
    fwi_loss(x, geometry, obs, misfit, direct_wave, mask, precond[, calc_grad]) -> (f, g float64[n], residuals)
    contract (fwi.py:236-246). The objective here is a stand-in quadratic with the same call signature, so the
    test needs no GPU; the GPU suite drives the real fwi_loss through the same contract."""
    import minimize as ref_minimize
    from optimize import NLCG    # (the reference's LBFGS core does `S==[]` on a memmap, which NumPy >= 2 rejects)
    from devito_fwi_b200.fwi import LazyResidual   # residuals may be lazy device arrays
    A = np.linspace(1., 4., 12)
    target = np.full(12, 0.25)
    calls = []

    def fake_fwi_loss(x, geometry, obs, misfit_func, direct_wave=None, mask=None, precond=True, calc_grad=True):
        calls.append(calc_grad)
        r = x - target
        return float(.5 * np.sum(A * r * r)), (A * r).astype(np.float64), [np.zeros((3, 2), np.float32)]

    ref_minimize.fwi_loss = fake_fwi_loss
    log = str(tmp_path / "log")
    opt = NLCG(ls_method='Bracket', step_len_init=0.5, max_ls=10, log_path=log, verbose=0)
    m = ref_minimize.minimize(opt, maxIter=8, ftol=1e-6, gtol=1e-8, log_path=log)
    x = m.run(np.ones(12), None, None, None, None, None, True, bounds=[0.01, 2.0])
    assert np.max(np.abs(x - target)) < 5e-2
    assert any(calls) and not all(calls)          # gradient evaluations and forward-only line-search trials
    assert os.path.exists(os.path.join(log, "misfit"))
    assert LazyResidual is not None


@pytest.mark.gpu
def test_reference_optimizer_drives_the_gpu_objective():
    """minimize.run + optimize.NLCG + misfit.least_square of the reference, unmodified, on the real fwi_loss for two
    iterations of the circle problem (scripts/run_reference_optimizer.py; a committed run: profiles/r02_reference_optimizer.txt).
    The reference checkout is not on the GPU box unless B2FWI_REFERENCE points at a copy."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("run_reference_optimizer",
                                                  os.path.join(ROOT, "scripts", "run_reference_optimizer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not os.path.isfile(os.path.join(mod.REF, "minimize.py")):
        pytest.skip("reference checkout not present")
    cwd = os.getcwd()
    try:
        out = mod.main(2, 3, 'NLCG')
    finally:
        os.chdir(cwd)
        for m in list(sys.modules):
            if m.split(".")[0] in ("fwi", "seismic", "devito", "w2", "minimize", "optimize", "misfit", "examples"):
                sys.modules.pop(m, None)
    h = out["objective_history"]
    assert len(h) >= 2 and h[-1] < h[0] and out["objective_after"] < h[0]
    assert out["gpu_launches"] > 0
