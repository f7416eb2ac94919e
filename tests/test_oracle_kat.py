"""Pins the CPU restatement (oracle/) against every known-answer value the reference
holds for the hot path. CPU only."""
import numpy as np
import pytest

from oracle import ref


# --------------------------------------------------------------------------------------
# seismic/inversion/fwi.py:13-121 -- circle model, 9 shots, so=4
def _circle_problem(dtype):
    shape, spacing, nbl = (101, 101), (10., 10.), 40

    def circle(vp_circle, vp_background, r=15):
        v = np.empty(shape, dtype=dtype)
        v[:] = vp_background
        a, b = shape[0] / 2, shape[1] / 2
        y, x = np.ogrid[-a:shape[0] - a, -b:shape[1] - b]
        v[x * x + y * y <= r * r] = vp_circle
        return v

    # demo_model's own space_order defaults to 2 (preset_models.py:43): it sets the CFL dt
    model = ref.RefModel((0., 0.), spacing, shape, 2, circle(3.0, 2.5), nbl=nbl, dtype=dtype)
    model0 = ref.RefModel((0., 0.), spacing, shape, 2, circle(2.5, 2.5), nbl=nbl, dtype=dtype)
    dt = float(model.critical_dt)
    nt, _, tv = ref.time_axis(0., 1000., dt)
    wav = ref.ricker(0.010, tv)
    rec = np.empty((101, 2))
    rec[:, 1] = np.linspace(0, model.domain_size[0], num=101)
    rec[:, 0] = 980.
    srcs = np.empty((9, 2), dtype=np.float32)
    srcs[:, 0] = 20.
    srcs[:, 1] = np.linspace(0., 1000, num=9)
    return model, model0, dt, nt, wav, rec, srcs


def _fwi_gradient(model, model0, dt, nt, wav, rec, srcs):
    grad = np.zeros(model.shape_pml, dtype=model.dtype)
    objective = 0.
    for i in range(srcs.shape[0]):
        d_obs, _ = ref.forward(model, srcs[i], rec, wav, nt, dt, space_order=4)
        d_syn, u0 = ref.forward(model0, srcs[i], rec, wav, nt, dt, save=True, space_order=4)
        residual = d_syn - d_obs
        objective += .5 * np.linalg.norm(residual.ravel()) ** 2
        ref.gradient(model0, residual, rec, u0, nt, dt, grad=grad, space_order=4)
    return objective, grad


def test_time_axis_and_dt_of_circle_kat():
    model, _, dt, nt, _, _, _ = _circle_problem(np.float32)
    assert np.isclose(dt, 2.041)
    assert nt == 491


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_circle_fwi_kat(dtype):
    prob = _circle_problem(dtype)
    model0 = prob[1]
    ff, update = _fwi_gradient(*prob)
    # seismic/inversion/fwi.py:95-97
    assert np.isclose(ff, 39113, atol=1e1, rtol=0)
    assert np.isclose(update.min(), -821, atol=1e1, rtol=0)
    assert np.isclose(update.max(), 2442, atol=1e1, rtol=0)
    if dtype is np.float64:
        # SURVEY Appendix C secondary pins
        assert np.isclose(ff, 39114.2274, rtol=1e-6)
        assert np.isclose(update.min(), -821.882, rtol=1e-5)
        assert np.isclose(update.max(), 2442.589, rtol=1e-5)
        # seismic/inversion/fwi.py:99-121: 5 gradient-descent iterations with box constraint
        history = []
        for _ in range(5):
            phi, direction = _fwi_gradient(*prob)
            history.append(phi)
            alpha = .05 / direction.max()
            model0.update_vp(np.clip(model0.vp + alpha * direction, 2.0, 3.5))
        assert np.isclose(history[-1], 3828, atol=1e1, rtol=0)
        assert np.allclose(history[1:], [24131.3645, 14098.3199, 7706.2682, 3828.7049], rtol=1e-5)


# --------------------------------------------------------------------------------------
# seismic/acoustic/acoustic_example.py:75-79 -- 3-D layered forward, fp64, |rec| = 459.1678
def test_3d_layers_forward_kat():
    shape, spacing, nbl, so = (50, 50, 50), (20., 20., 20.), 40, 4
    v = np.empty(shape, dtype=np.float64)
    v[:] = 1.5
    vp_i = np.linspace(1.5, 3.5, 3)
    for i in range(1, 3):
        v[..., i * int(shape[-1] / 3):] = vp_i[i]
    model = ref.RefModel((0., 0., 0.), spacing, shape, so, v, nbl=nbl, dtype=np.float64)
    dt = float(model.critical_dt)
    nt, _, tv = ref.time_axis(0., 1000., dt)
    # seismic/utils.py:12-47 setup_geometry / setup_rec_coords
    src = np.array(model.domain_size) * .5
    src[-1] = 0. + spacing[-1]
    recx = np.linspace(0., model.domain_size[0], shape[0])
    recy = np.linspace(0., model.domain_size[1], shape[1])
    rec = np.empty((shape[0] * shape[1], 3))
    rec[:, 0] = np.repeat(recx, shape[1])
    rec[:, 1] = np.tile(recy, shape[0])
    rec[:, 2] = 2 * spacing[-1]
    d, _ = ref.forward(model, src, rec, ref.ricker(0.010, tv), nt, dt)
    assert np.isclose(np.linalg.norm(d.ravel()), 459.1678, rtol=1e-3, atol=0)
    assert np.isclose(np.linalg.norm(d.ravel()), 459.3391, rtol=1e-5)


# --------------------------------------------------------------------------------------
# seismic/acoustic/accuracy.ipynb cells 5-16 -- constant medium vs analytic Hankel solution
def test_accuracy_notebook_kat():
    from scipy.special import hankel2
    nt, dt, f0, c0 = 1501, 0.1, .09, 1.5
    model = ref.RefModel((0., 0.), (.5, .5), (801, 801), 20, c0, nbl=40, dtype=np.float64, dt=dt)
    num, _, tv = ref.time_axis(0., dt * (nt - 1), dt)
    assert num == nt
    wav = ref.ricker(f0, tv, t0=1.5 / f0)
    d, _ = ref.forward(model, [200., 200.], [260., 260.], wav, nt, dt, space_order=8)
    # cell 14 output
    assert "%+.6e" % d.min() in ("-5.349877e-03", "-5.349875e-03", "-5.349876e-03")
    assert np.isclose(d.min(), -5.349877e-03, rtol=1e-6)
    assert np.isclose(d.max(), +8.529867e-03, rtol=1e-6)

    # cells 12-13: analytic solution
    def ricker(f, T, dt_, t0):
        t = np.linspace(-t0, T - t0, int(T / dt_))
        tt = (np.pi ** 2) * (f ** 2) * (t ** 2)
        return (1.0 - 2.0 * tt) * np.exp(- tt)

    def analytical(nt_, time, dt_):
        nf = int(nt_ / 2 + 1)
        df = 1.0 / time[-1]
        faxis = df * np.arange(nf)
        R = np.fft.fft(ricker(f0, time[-1], dt_, 1.5 / f0))[0:nf]
        U_a = np.zeros((nf), dtype=complex)
        a = np.arange(1, nf - 1)
        k = 2 * np.pi * faxis[a] / c0
        U_a[a] = -1j * np.pi * hankel2(0.0, k * np.sqrt(60. ** 2 + 60. ** 2)) * R[a]
        U_t = 1.0 / (2.0 * np.pi) * np.real(np.fft.ifft(U_a[:], nt_))
        return np.real(U_t) * (.5 ** 2)

    time1 = np.linspace(0.0, 3000., 30001)
    U_t = analytical(30001, time1, time1[1] - time1[0])[0:1501]
    err = np.linalg.norm(U_t[:-1] - d[:-1, 0], 2) / np.sqrt(nt)
    assert np.isclose(err, 1.1265077536675204e-05, rtol=1e-3)      # cell 16 output


def test_3d_layers_forward_free_surface_kat():
    """seismic/acoustic/acoustic_example.py:75-79, fs=True, float32: |rec|_2 = 369.955 (rtol 1e-3) - pins the
    mirrored top rows (operators.py:8-35) and the one-sided sponge / padding (model.py:33,102-109,157)."""
    shape, spacing, nbl, so = (50, 50, 50), (20., 20., 20.), 40, 4
    v = np.empty(shape, dtype=np.float32)
    v[:] = 1.5
    vp_i = np.linspace(1.5, 3.5, 3)
    for i in range(1, 3):
        v[..., i * int(shape[-1] / 3):] = vp_i[i]
    norms = {}
    for dtype in (np.float32, np.float64):
        model = ref.RefModel((0., 0., 0.), spacing, shape, so, v.astype(dtype), nbl=nbl, dtype=dtype, fs=True)
        assert model.shape_pml == (130, 130, 90) and float(model.origin_pml[2]) == 0.0
        assert model.damp[65, 65, 0] == 0 and model.damp[65, 65, -1] > 0 and model.damp[0, 65, 0] > 0
        dt = float(model.critical_dt)
        nt, _, tv = ref.time_axis(0., 1000., dt)
        src = np.array(model.domain_size) * .5
        src[-1] = 0. + spacing[-1]
        recx = np.linspace(0., model.domain_size[0], shape[0])
        recy = np.linspace(0., model.domain_size[1], shape[1])
        rec = np.empty((shape[0] * shape[1], 3))
        rec[:, 0] = np.repeat(recx, shape[1])
        rec[:, 1] = np.tile(recy, shape[0])
        rec[:, 2] = 2 * spacing[-1]
        d, _ = ref.forward(model, src, rec, ref.ricker(0.010, tv), nt, dt)
        norms[dtype] = float(np.linalg.norm(d.astype(np.float64).ravel()))
    print("free-surface KAT: |rec| fp32 %.4f fp64 %.4f (reference 369.955)" % (norms[np.float32], norms[np.float64]))
    assert np.isclose(norms[np.float32], 369.955, rtol=1e-3, atol=0)
    assert np.isclose(norms[np.float64], norms[np.float32], rtol=1e-4)
