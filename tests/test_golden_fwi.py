"""Golden vectors of the reference's own fwi.py glue (tests/golden/make_fwi_golden.py: the UNMODIFIED reference
fwi_obj_multi / fwi_loss / fix_source_illumination + misfit.least_square, with the pinned oracle as propagator).
  * CPU: the oracle's numpy restatement of that glue (oracle/ref.py) must reproduce them;
  * GPU: the CUDA path (resident and streaming engines) must reproduce them within the stated fp32 tolerances."""
import os

import numpy as np
import pytest

from tests.golden.make_fwi_golden import problem, geometries
from tests.util import ref_model, rel_l2
from oracle import ref

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fwi_obj_small.npz"))


def test_oracle_glue_reproduces_reference_fwi_py():
    p = problem()
    g_true, g_init, g_const = geometries(p)
    rm_true, rm_init, rm_const = (ref_model(g.model) for g in (g_true, g_init, g_const))
    nt, dt = g_init.nt, float(g_init.dt)
    wav = np.float64(g_init.src.data[:, :1])
    fw = lambda rm: [np.float32(ref.forward(rm, p["src"][i], p["rec"], wav, nt, dt)[0]) for i in range(3)]  # noqa: E731
    obs, dw = fw(rm_true), fw(rm_const)
    assert rel_l2(obs[0], GOLD["obs0"]) < 1e-7
    f, g, res = ref.fwi_obj_multi(rm_init, p["src"], p["rec"], wav, nt, dt, obs, direct_wave=dw, mask=p["mask"],
                                  precond=True, calc_grad=True)
    assert np.isclose(f, GOLD["f_full"], rtol=1e-5)
    assert rel_l2(g, GOLD["g_full"]) < 1e-5
    f, g, _ = ref.fwi_obj_multi(rm_init, p["src"], p["rec"], wav, nt, dt, obs, precond=False, calc_grad=True)
    assert np.isclose(f, GOLD["f_plain"], rtol=1e-5)
    assert rel_l2(g, GOLD["g_plain"]) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["auto", "stream"])
def test_cuda_objective_reproduces_reference_fwi_py(engine):
    from devito_fwi_b200 import fwi
    p = problem()
    g_true, g_init, g_const = geometries(p)
    fwi.ENGINE = engine
    try:
        obs = fwi.fm_multi(g_true)
        dw = fwi.fm_multi(g_const)
        assert rel_l2(obs[0].data, GOLD["obs0"]) <= 1e-5
        f, g, res = fwi.fwi_obj_multi(g_init, obs, fwi.least_square, dw, p["mask"], True, True)
        print("%s engine vs reference fwi.py: f %.2e  g %.2e" % (engine, abs(f - GOLD["f_full"]) / GOLD["f_full"],
                                                                rel_l2(g, GOLD["g_full"])))
        assert abs(f - GOLD["f_full"]) / GOLD["f_full"] <= 1e-4
        assert rel_l2(g, GOLD["g_full"]) <= 1e-4
        assert rel_l2(np.asarray(res[0]), GOLD["res0_full"]) <= 1e-4
        f, g, _ = fwi.fwi_obj_multi(g_init, obs, fwi.least_square, None, None, False, True)
        assert abs(f - GOLD["f_plain"]) / GOLD["f_plain"] <= 1e-4 and rel_l2(g, GOLD["g_plain"]) <= 1e-4
        x = (1. / (np.float64(p["vp_init"]) ** 2)).ravel()
        f, g, _ = fwi.fwi_loss(x, g_init, obs, fwi.least_square, dw, p["mask"], True, True)
        assert abs(f - GOLD["f_loss"]) / GOLD["f_loss"] <= 1e-4 and rel_l2(g, GOLD["g_loss"]) <= 1e-4
    finally:
        fwi.ENGINE = 'auto'


@pytest.mark.gpu
def test_fwi_obj_single_sums_to_reference_objective():
    """fwi_obj_single (fwi.py:131-173, per-shot API on the streaming engine): the shot sum of its (f, crop_grad, illum)
    reproduces the reference's fwi_obj_multi golden values."""
    from devito_fwi_b200 import fwi
    p = problem()
    g_true, g_init, g_const = geometries(p)
    obs = fwi.fm_multi(g_true)
    f_sum, g_sum, il_sum = 0., 0., 0.
    for i in range(3):
        f, g, res, il = fwi.fwi_obj_single(fwi._shot_geometry(g_init, i), obs[i], fwi.least_square, None,
                                           g_init.dt, True)
        assert g.shape == g_init.model.shape and il.shape == g.shape and res.shape == (g_init.nt, 13)
        f_sum, g_sum, il_sum = f_sum + f, g_sum + g, il_sum + il
    assert abs(f_sum - GOLD["f_plain"]) / GOLD["f_plain"] <= 1e-4
    assert rel_l2(g_sum.ravel(), GOLD["g_plain"]) <= 1e-4
    g_pre = (g_sum / np.sqrt(il_sum + 1e-30)) * p["mask"]
    assert rel_l2(g_pre.ravel(), GOLD["g_full"]) <= 1e-4
    f0, g0, r0, il0 = fwi.fwi_obj_single(fwi._shot_geometry(g_init, 0), obs[0], fwi.least_square, None, None, False)
    assert g0 is None and il0 is None and f0 > 0
