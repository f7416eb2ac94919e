"""GPU parity of the SM-resident 2-D engine (one cluster per shot, all shots in one launch) against
the fp64 oracle and against the per-step streaming engine, on the named 2-D configurations."""
import numpy as np
import pytest

from tests.util import ref_model, rel_l2
from oracle import ref

pytestmark = pytest.mark.gpu

TOL_TRACE = 1e-5
TOL_GRAD = 1e-4


def _oracle_shot(geom_true, geom_init, i):
    rm_true, rm_init = ref_model(geom_true.model), ref_model(geom_init.model)
    nt, dt = geom_init.nt, float(geom_init.dt)
    wav = np.float64(geom_init.src.data[:, :1])
    o64, _ = ref.forward(rm_true, geom_true.src_positions[i], geom_true.rec_positions, wav, nt, dt)
    s64, u64 = ref.forward(rm_init, geom_init.src_positions[i], geom_init.rec_positions, wav, nt, dt, save=True)
    g64 = ref.gradient(rm_init, s64 - o64, geom_init.rec_positions, u64, nt, dt)
    illum = (u64 * u64).sum(axis=0)
    return o64, s64, g64, illum


@pytest.mark.parametrize("config", ["marmousi", "circle6", "circle4", "marmousi2"])
def test_resident_engine_vs_oracle(config):
    import torch
    import devito_fwi_b200 as b
    from devito_fwi_b200 import configs
    from devito_fwi_b200.resident import ResidentSurvey
    if config == "marmousi":
        g_true, g_init = configs.marmousi()[:2]
        shots = [0, 14, 28]
    elif config == "marmousi2":
        g_true, g_init = configs.marmousi2()[:2]
        shots = [7]
    else:
        g_true, g_init = configs.circle(space_order=6 if config == "circle6" else 4)
        shots = [0, 5]
    assert ResidentSurvey.supported(g_init)
    sv_true = ResidentSurvey(g_true, shots)
    sv = ResidentSurvey(g_init, shots)
    obs = sv_true.forward().clone()
    syn = sv.forward(save=True, illum=True).clone()
    res = (syn - obs).contiguous()
    grad = sv.crop(sv.gradient(res)).cpu().numpy()
    illum = sv.crop(sv.illum).cpu().numpy()
    # determinism: a second run is bitwise identical
    syn2 = sv.forward(save=True, illum=True)
    assert torch.equal(syn, syn2)
    grad2 = sv.crop(sv.gradient(res)).cpu().numpy()
    assert np.array_equal(grad, grad2)
    nbl = g_init.model.nbl
    for k, i in enumerate(shots):
        o64, s64, g64, il64 = _oracle_shot(g_true, g_init, i)
        et, eo = rel_l2(syn[k].cpu().numpy(), s64), rel_l2(obs[k].cpu().numpy(), o64)
        # gradient driven by the oracle's own residual, to separate it from the trace error
        res64 = torch.from_numpy(np.float32(s64 - o64)[None]).cuda()
        print("%s shot %d (cluster=%d): traces %.2e / %.2e  illum %.2e" % (
            config, i, sv.plan.cluster, et, eo, rel_l2(illum[k], il64[nbl:-nbl, nbl:-nbl])))
        assert et <= TOL_TRACE and eo <= TOL_TRACE
        assert rel_l2(illum[k], il64[nbl:-nbl, nbl:-nbl]) <= TOL_GRAD
        eg = rel_l2(grad[k], g64[nbl:-nbl, nbl:-nbl])
        print("   gradient (own residual) %.2e" % eg)
        assert eg <= TOL_GRAD           # (measured 1e-6 .. 6e-6 on all four named configurations)
    # same residual on both sides -> the gradient tolerance proper
    o64, s64, g64, _ = _oracle_shot(g_true, g_init, shots[0])
    res_all = res.clone()
    res_all[0] = torch.from_numpy(np.float32(s64 - o64)).cuda()
    g0 = sv.crop(sv.gradient(res_all.contiguous()))[0].cpu().numpy()
    eg = rel_l2(g0, g64[nbl:-nbl, nbl:-nbl])
    print("   gradient (oracle residual) %.2e" % eg)
    assert eg <= TOL_GRAD


def test_objective_resident_vs_streaming_vs_oracle():
    """fwi_obj_multi / fwi_loss on Marmousi (5 shots, direct-wave subtraction, bathymetry mask,
    illumination preconditioning): resident engine + on-device L2 misfit == streaming engine + host
    plug-in misfit == numpy restatement of fwi.py."""
    import devito_fwi_b200 as b
    from devito_fwi_b200 import configs, fwi
    g_true, g_init, g_const, mask = configs.marmousi(nsrc=5)
    obs = fwi.fm_multi(g_true)
    dw = fwi.fm_multi(g_const)
    x = 1. / (g_init.model.vp.data[40:-40, 40:-40].astype(np.float64) ** 2)
    f_r, g_r, res_r = fwi.fwi_loss(x.ravel(), g_init, obs, fwi.least_square, dw, mask, True, True)
    assert len(res_r) == 5 and np.asarray(res_r[0]).shape == (g_init.nt, 300)

    def host_misfit(syn, o):       # a plug-in the engine cannot recognise -> host round trip
        r = syn - o
        return .5 * np.linalg.norm(r.flatten()) ** 2, r
    f_h, g_h, _ = fwi.fwi_loss(x.ravel(), g_init, obs, host_misfit, dw, mask, True, True)
    fwi.ENGINE = 'stream'
    try:
        f_s, g_s, _ = fwi.fwi_loss(x.ravel(), g_init, obs, fwi.least_square, dw, mask, True, True)
    finally:
        fwi.ENGINE = 'auto'
    print("resident vs host-misfit: f %.2e g %.2e | resident vs streaming: f %.2e g %.2e" % (
        abs(f_r - f_h) / f_h, rel_l2(g_r, g_h), abs(f_r - f_s) / f_s, rel_l2(g_r, g_s)))
    assert abs(f_r - f_h) / f_h < 1e-5 and rel_l2(g_r, g_h) < 1e-6   # host misfit sums in fp32
    assert abs(f_r - f_s) / f_s < 1e-5 and rel_l2(g_r, g_s) < TOL_GRAD    # measured 3e-6 / 4e-6

    rm_true, rm_init, rm_const = (ref_model(g.model) for g in (g_true, g_init, g_const))
    nt, dt = g_init.nt, float(g_init.dt)
    wav = np.float64(g_init.src.data[:, :1])
    fw = lambda rm: [ref.forward(rm, g_init.src_positions[i], g_init.rec_positions, wav, nt, dt)[0]  # noqa: E731
                     for i in range(5)]
    f64, g64, _ = ref.fwi_obj_multi(rm_init, g_init.src_positions, g_init.rec_positions, wav, nt, dt,
                                    fw(rm_true), direct_wave=fw(rm_const), mask=mask, precond=True,
                                    calc_grad=True)
    print("resident vs oracle: f %.2e g %.2e" % (abs(f_r - f64) / f64, rel_l2(g_r, g64)))
    assert abs(f_r - f64) / f64 <= 1e-5
    assert rel_l2(g_r, g64) <= TOL_GRAD         # measured 2e-7 / 2e-6
    # forward-only evaluation of the line search
    f2, g2, _ = fwi.fwi_loss(x.ravel(), g_init, obs, fwi.least_square, dw, mask, True, False)
    assert np.isclose(f2, f_r, rtol=1e-9) and not g2.any()


def test_objective_in_launch_groups_matches_single_group(monkeypatch):
    """Surveys with more shots than one wave of clusters are split into launch groups with their own decomposition
    (resident.partition_shots). A forced split (3 shots on 4-CTA clusters with 12-row strips = the long-strip
    kernel, 2 shots on 16-CTA clusters with 3-row strips = the short-strip kernel) must give the objective,
    gradient, residuals and fm_multi records of the single-group run."""
    from devito_fwi_b200 import configs, fwi, resident
    g_true, g_init, g_const, mask = configs.marmousi(nsrc=5)
    x = (1. / (g_init.model.vp.data[40:-40, 40:-40].astype(np.float64) ** 2)).ravel()
    groups = resident.partition_shots(g_init.model.grid, 8, 40, 5)
    assert groups == [(5, groups[0][1])]                      # five shots: one wave, one group
    assert sum(k for k, _ in resident.partition_shots(g_init.model.grid, 8, 40, 300)) == 300
    obs, dw = fwi.fm_multi(g_true), fwi.fm_multi(g_const)
    f1, g1, r1 = fwi.fwi_loss(x, g_init, obs, fwi.least_square, dw, mask, True, True)
    r1 = [np.asarray(r).copy() for r in r1]
    fwi._SURVEYS.clear()
    grid = g_init.model.grid
    forced = [(3, resident.plan_exact(grid, 8, 40, 4, 12)), (2, resident.plan_exact(grid, 8, 40, 16, 3))]
    assert all(p is not None for _, p in forced)
    monkeypatch.setattr(resident, 'partition_shots', lambda *a, **k: forced)
    svs = fwi._resident_surveys(g_init, list(range(5)))
    assert [sv.nshots for sv in svs] == [3, 2] and [sv.plan.cluster for sv in svs] == [4, 16]
    assert [sv.shots for sv in svs] == [[0, 1, 2], [3, 4]]
    f2, g2, r2 = fwi.fwi_loss(x, g_init, obs, fwi.least_square, dw, mask, True, True)
    obs2 = fwi.fm_multi(g_true)
    fwi._SURVEYS.clear()
    assert abs(f2 - f1) <= 1e-6 * abs(f1) and rel_l2(g2, g1) < 1e-5
    assert len(r2) == 5
    for k in range(5):
        assert rel_l2(np.asarray(r2[k]), r1[k]) < 1e-5
        assert rel_l2(obs2[k].data, obs[k].data) < 1e-5


def test_resident_many_shots_small_grid_vs_streaming():
    """More shots than resident clusters (several waves), a 2-CTA cluster, sources / receivers in the
    sponge and outside the grid: the batched resident engine must agree with the per-shot streaming
    engine (two independent fp32 implementations, both checked against the oracle elsewhere)."""
    import torch
    import devito_fwi_b200 as b
    from devito_fwi_b200 import fwi
    from devito_fwi_b200.resident import ResidentSurvey
    shape, nbl, so = (120, 90), 20, 8
    vp = np.full(shape, 1.8, dtype=np.float32)
    vp[:, 45:] = 2.6
    vp[40:70, 20:40] = 3.1
    model = b.Model(origin=(0., 0.), spacing=(10., 10.), shape=shape, space_order=so, vp=vp, nbl=nbl, dt=1.2)
    nsrc = 40
    src = np.stack([np.linspace(-50., 1240., nsrc), np.full(nsrc, 25.)], axis=1)      # first / last in the sponge
    rec = np.stack([np.linspace(-300., 1500., 37), np.full(37, 33.3)], axis=1)        # first / last outside the grid
    geom = b.AcquisitionGeometry(model, rec, src, 0., 420., f0=0.02, src_type='Ricker')
    sv = ResidentSurvey(geom, min_cluster=2)          # a 2-CTA cluster: exercises the DSMEM halo exchange on both sides
    assert sv.plan.cluster == 2
    from devito_fwi_b200.resident import choose_plan
    assert choose_plan(model.grid, so, nbl, nsrc).cluster >= 2
    syn = sv.forward(save=True, illum=True).clone()
    g_res = sv.crop(sv.gradient(syn.contiguous())).cpu().numpy()
    il_res = sv.crop(sv.illum).cpu().numpy()
    fwi.ENGINE = 'stream'
    try:
        for i in (0, 1, 17, 39):
            gi = fwi._shot_geometry(geom, i)
            solver = b.AcousticWaveSolver(model, gi, space_order=so)
            illum = b.Function(name='illum', grid=model.grid)
            d, u, _ = solver.forward(save=True, illum=illum)
            assert rel_l2(syn[i].cpu().numpy(), d.data) < 2e-5
            res = b.Receiver(name='r', grid=model.grid, time_range=gi.time_axis, coordinates=rec)
            res.data[:] = syn[i].cpu().numpy()
            grad, _ = solver.gradient(rec=res, u=u)
            assert rel_l2(g_res[i], grad.data[nbl:-nbl, nbl:-nbl]) < 1e-4
            assert rel_l2(il_res[i], illum.data[nbl:-nbl, nbl:-nbl]) < 1e-4
    finally:
        fwi.ENGINE = 'auto'
    assert not np.any(syn[:, :, 0].cpu().numpy()) and not np.any(syn[:, :, -1].cpu().numpy())   # outside the grid


def test_fm_multi_and_host_misfit_plugin_roundtrip():
    """fm_multi returns Receivers whose .data materialises lazily; a numpy plug-in misfit (1-D style) works."""
    import devito_fwi_b200 as b
    from devito_fwi_b200 import configs, fwi
    g_true, g_init = configs.circle(space_order=4, nsrc=2)
    obs = fwi.fm_multi(g_true)
    assert len(obs) == 2 and obs[0].data.shape == (g_true.nt, 201) and obs[0].data.dtype == np.float32
    assert np.abs(obs[0].data).max() > 0 and not np.any(obs[0].data[0]) and not np.any(obs[0].data[-1])

    def trace_normalised(syn, o):          # any (syn, obs) -> (f, adjoint source) callable is accepted
        r = syn - o
        return float(.5 * np.sum(r.astype(np.float64) ** 2)), r
    f, g, res = fwi.fwi_obj_multi(g_init, obs, trace_normalised, None, None, False, True)
    f2, g2, _ = fwi.fwi_obj_multi(g_init, obs, fwi.least_square, None, None, False, True)
    assert np.isclose(f, f2, rtol=1e-6) and rel_l2(g, g2) < 1e-6
    assert isinstance(res[0], np.ndarray) and res[0].shape == (g_init.nt, 201)


def test_gradient_taylor_and_linearity():
    """Property tests through the product path (patterns of seismic/self_adjoint/test_wavesolver_iso.py):
    linearity F(a s) = a F(s), and the Taylor test of the objective/gradient pair
    f(m + h dm) - f(m) - h <g, dm> = O(h^2) with the un-preconditioned, un-masked gradient (fwi.py:236-246)."""
    import devito_fwi_b200 as b
    from devito_fwi_b200 import configs, fwi
    g_true, g_init = configs.circle(space_order=4, nsrc=2)
    obs = fwi.fm_multi(g_true)
    # linearity in the source amplitude (resident engine, batched shots)
    from devito_fwi_b200.resident import ResidentSurvey
    sv = ResidentSurvey(g_init, [0, 1])
    d1 = sv.forward().clone()
    # Scaling by a power of two is exact in IEEE arithmetic EXCEPT in the denormal range, which the numerical
    # precursor ahead of the wavefront crosses (1e-44 .. 1e-38); one differing bit there decorrelates all later
    # roundings, so the two runs are two fp32 realisations that differ by the round-off noise floor (~1e-6).
    sv.src.mul_(2.0)
    d2 = sv.forward().clone()
    assert rel_l2(d2.cpu().numpy(), 2.0 * d1.cpu().numpy()) < 1e-5
    sv.src.mul_(1.5)
    d3 = sv.forward().clone()
    assert rel_l2(d3.cpu().numpy(), 3.0 * d1.cpu().numpy()) < 1e-5

    shape = g_init.model.shape
    m0 = (1. / (g_init.model.vp.data[40:-40, 40:-40].astype(np.float64) ** 2)).ravel()
    f0, g0, _ = fwi.fwi_loss(m0, g_init, obs, fwi.least_square, None, None, False, True)
    rng = np.random.default_rng(3)
    xx, zz = np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), indexing='ij')
    dm = (1e-3 * np.exp(-((xx - 100) ** 2 + (zz - 90) ** 2) / 800.)).ravel()     # smooth slowness bump
    errs = []
    hs = [1.0, 0.5, 0.25]
    for h in hs:
        fh, _, _ = fwi.fwi_loss(m0 + h * dm, g_init, obs, fwi.least_square, None, None, False, False)
        errs.append((abs(fh - f0), abs(fh - f0 - h * np.dot(g0, dm))))
    print("Taylor test: |f(m+h dm)-f(m)|, |... - h<g,dm>|:", errs)
    # first-order remainder halves, second-order remainder quarters (within fp32 noise)
    for k in range(len(hs) - 1):
        assert 1.7 < errs[k][0] / errs[k + 1][0] < 2.3
        assert errs[k][1] / errs[k + 1][1] > 3.3
    assert errs[-1][1] < 0.05 * errs[-1][0]


def test_marmousi_fwi_example_reduces_the_objective():
    """A few L-BFGS iterations of the Marmousi L2 FWI driver (examples/marmousi_fwi.py) on 5 shots."""
    import sys
    import examples.marmousi_fwi as ex
    argv = sys.argv
    sys.argv = ["marmousi_fwi.py", "--nsrc", "5", "--maxiter", "3", "--odir", "/tmp/b2fwi_result"]
    try:
        history = ex.main()
    finally:
        sys.argv = argv
    assert len(history) >= 3
    assert min(history) < 0.8 * history[0]


def test_observed_data_cache_follows_host_modifications():
    """The device copy of the observed records is re-used between evaluations, but any access to a record's
    host view (obs[i].data - the caller may have changed it) forces a fresh upload."""
    from devito_fwi_b200 import configs, fwi
    g_true, g_init = configs.circle(space_order=4, nsrc=2)
    obs = fwi.fm_multi(g_true)
    f1, _, _ = fwi.fwi_obj_multi(g_init, obs, fwi.least_square, None, None, False, False)
    f1b, _, _ = fwi.fwi_obj_multi(g_init, obs, fwi.least_square, None, None, False, False)
    assert f1 == f1b
    obs[0].data[:] = 0.0                         # in-place host modification
    f2, _, _ = fwi.fwi_obj_multi(g_init, obs, fwi.least_square, None, None, False, False)
    assert f2 != f1
    obs2 = fwi.fm_multi(g_true)                  # a different list object with the original data
    f3, _, _ = fwi.fwi_obj_multi(g_init, obs2, fwi.least_square, None, None, False, False)
    assert np.isclose(f3, f1, rtol=1e-12)


def test_objective_with_observed_data_on_another_time_axis():
    """fwi.py:47-57 / source.py:140-170: observed data recorded at another sampling step are
    spline-resampled onto the modelling axis inside the objective (the per-shot streaming branch, not the batched
    one). Up-sampling the records to dt/2 keeps every original sample as a spline node, so the objective and gradient
    must come out the same as with the original records."""
    from devito_fwi_b200 import configs, fwi
    g_true, g_init, g_const, mask = configs.marmousi(nsrc=2, tn=1200.)
    x = (1. / (g_init.model.vp.data[40:-40, 40:-40].astype(np.float64) ** 2)).ravel()
    obs, dw = fwi.fm_multi(g_true), fwi.fm_multi(g_const)
    f1, g1, _ = fwi.fwi_loss(x, g_init, obs, fwi.least_square, dw, mask, True, True)
    dt = float(g_init.dt)
    obs_fine = [o.resample(dt=dt / 2) for o in obs]
    assert obs_fine[0].data.shape[0] == 2 * (g_init.nt - 1) + 1
    f2, g2, res2 = fwi.fwi_loss(x, g_init, obs_fine, fwi.least_square, dw, mask, True, True)
    assert len(res2) == 2 and np.asarray(res2[0]).shape == (g_init.nt, 300)
    print("resampled observed data: f %.3e g %.3e" % (abs(f2 - f1) / f1, rel_l2(g2, g1)))
    assert abs(f2 - f1) <= 1e-5 * f1 and rel_l2(g2, g1) <= TOL_GRAD     # measured 1e-6 / 2e-6


def test_residuals_are_snapshots_and_line_search_evaluations_return_fval_only():
    """The residuals of one evaluation must survive later evaluations on the same cached survey (the reference returns
    independent arrays; a line search evaluates the objective again before anyone looks at them), and a
    calc_grad=False evaluation returns the objective with a zero gradient (minimize.py:59-86)."""
    from devito_fwi_b200 import configs, fwi
    g_true, g_init = configs.circle(space_order=4, nsrc=2)
    obs = fwi.fm_multi(g_true)
    nbl = g_init.model.nbl
    x0 = (1. / (g_init.model.vp.data[nbl:-nbl, nbl:-nbl].astype(np.float64) ** 2)).ravel()
    f1, g1, r1 = fwi.fwi_loss(x0, g_init, obs, fwi.least_square)
    f2, g2, r2 = fwi.fwi_loss(x0 * 1.03, g_init, obs, fwi.least_square, calc_grad=False)
    assert f2 != f1 and not g2.any() and g1.any()
    first = [np.asarray(r).copy() for r in r1]          # read only now, after the second evaluation
    f3, g3, r3 = fwi.fwi_loss(x0, g_init, obs, fwi.least_square)
    assert f3 == f1 and np.array_equal(g3, g1)
    for a, b in zip(first, r3):
        assert np.array_equal(a, np.asarray(b))
    assert not np.array_equal(first[0], np.asarray(r2[0]))
    assert float(0.5 * sum(np.sum(np.float64(a) ** 2) for a in first)) == pytest.approx(f1, rel=1e-6)
