"""GPU parity of the on-device 2-D quadratic-Wasserstein misfit (b2fwi_qw2d_misfit) against the reference's own
back-and-forth solver: golden vectors (tests/golden/qw2d_small.npz) and, where the compiled reference is at hand
(oracle/_ref/libqw2d_ref.so travels with the repository snapshot), live comparisons at larger sizes.

Tolerances. The solver is a fixed number of single-precision fixed-point iterations whose step size is steered by
threshold tests; the two implementations differ in the DCT (FFTW stand-in vs cuFFT) and in the accumulation order of
the push-forward, i.e. at the 1e-7 level per operation. Measured agreement: loss to ~1e-5 relative, adjoint source
to ~1e-4 relative L2; asserted: 1e-4 and 2e-3."""
import os

import numpy as np
import pytest

from tests.util import rel_l2
from tests.golden.make_qw2d_golden import records

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qw2d_small.npz")
TOL_LOSS, TOL_ADJ = 1e-4, 2e-3


def _device_qw2d(f_list, g_list, gamma, steps, scale, dw=None):
    import ctypes
    import torch
    from devito_fwi_b200 import _lib
    lib = _lib.lib()
    syn = torch.from_numpy(np.stack(f_list)).cuda()
    obs = torch.from_numpy(np.stack(g_list)).cuda()
    dwd = torch.from_numpy(np.stack(dw)).cuda() if dw is not None else None
    ns, nt, nrec = syn.shape
    adj = torch.empty_like(syn)
    fval = torch.zeros(1, dtype=torch.float64, device='cuda')
    loss = torch.zeros(ns, dtype=torch.float32, device='cuda')
    scratch = torch.empty(int(lib.b2fwi_qw2d_scratch_bytes(nt, nrec, ns)), dtype=torch.uint8, device='cuda')
    _lib.check(lib.b2fwi_qw2d_misfit(syn.data_ptr(), obs.data_ptr(), dwd.data_ptr() if dwd is not None else None, nt, nrec,
                                     ns, ctypes.c_double(gamma), steps, ctypes.c_float(scale), adj.data_ptr(),
                                     fval.data_ptr(), loss.data_ptr(), scratch.data_ptr(), None))
    torch.cuda.synchronize()
    return loss.cpu().numpy(), adj.cpu().numpy(), float(fval.item())


@pytest.mark.parametrize("name", ["a", "b"])
def test_qw2d_matches_reference_golden(name):
    g = np.load(GOLD)
    steps, scale, gamma = g[name + "_par"]
    loss, adj, fval = _device_qw2d([g[name + "_f"]], [g[name + "_g"]], float(gamma), int(steps), float(scale))
    el, ea = abs(loss[0] - g[name + "_loss"]) / abs(g[name + "_loss"]), rel_l2(adj[0], g[name + "_adj"])
    print("QW2D golden %s: loss %.8e vs %.8e (rel %.2e), adjoint source rel-L2 %.2e" % (name, loss[0], g[name + "_loss"], el, ea))
    assert el <= TOL_LOSS and ea <= TOL_ADJ
    assert abs(fval - loss[0]) <= 1e-7 * abs(fval)


def test_qw2d_batch_equals_single_records_and_direct_wave():
    """Records of a batch are independent (bitwise the single-record results), the result is repeatable from call to call
    (fixed-point accumulation in the push-forward), and the direct-wave subtraction is (syn - dw), (obs - dw)."""
    pairs = [records(80, 37, s) for s in (1, 2, 3)]
    fs, gs = [p[0] for p in pairs], [p[1] for p in pairs]
    loss_b, adj_b, fval_b = _device_qw2d(fs, gs, 1.01, 5, 4.0)
    loss_b2, adj_b2, _ = _device_qw2d(fs, gs, 1.01, 5, 4.0)
    assert np.array_equal(adj_b, adj_b2) and np.array_equal(loss_b, loss_b2)
    for k in range(3):
        loss_1, adj_1, _ = _device_qw2d([fs[k]], [gs[k]], 1.01, 5, 4.0)
        assert loss_1[0] == loss_b[k] and np.array_equal(adj_1[0], adj_b[k])
    assert abs(fval_b - float(np.sum(np.float64(loss_b)))) <= 1e-12
    dw = [0.3 * fs[0]] * 3
    loss_d, adj_d, _ = _device_qw2d([f + d for f, d in zip(fs, dw)], [g + d for g, d in zip(gs, dw)], 1.01, 5, 4.0, dw=dw)
    f2 = [(f + d) - d for f, d in zip(fs, dw)]
    g2 = [(g + d) - d for g, d in zip(gs, dw)]
    loss_e, adj_e, _ = _device_qw2d(f2, g2, 1.01, 5, 4.0)
    assert np.array_equal(loss_d, loss_e) and np.array_equal(adj_d, adj_e)


@pytest.mark.parametrize("nt,nrec,steps", [(240, 64, 15), (1501, 340, 15)])
def test_qw2d_vs_compiled_reference(nt, nrec, steps):
    """Live against the reference's fot2d.c (oracle/_ref), up to the Marmousi2 record size (marmousi2_fwi.py:
    nt = 1501, 340 receivers, gamma = 1.01, num_steps = 15, step_scale = 4)."""
    from oracle import ref_qw2d
    if not ref_qw2d.available():
        pytest.skip("compiled reference (oracle/_ref/libqw2d_ref.so) not present")
    f, g = records(nt, nrec, 7)
    loss_r, adj_r = ref_qw2d.qwasserstein_2d(f, g, 1.01, steps, 4.0)
    loss, adj, _ = _device_qw2d([f], [g], 1.01, steps, 4.0)
    el, ea = abs(loss[0] - loss_r) / abs(loss_r), rel_l2(adj[0], adj_r)
    print("QW2D %dx%d, %d steps: loss %.8e vs %.8e (rel %.2e), adjoint source rel-L2 %.2e" % (nt, nrec, steps, loss[0], loss_r, el, ea))
    assert el <= TOL_LOSS and ea <= TOL_ADJ


def test_fwi_loss_with_qw2d_on_device_vs_host_plugin():
    """fwi_loss with qWasserstein(method='2d'): the device path (all shots in one call, residual injected straight from
    HBM) against the same objective with the reference's solver as a HOST plug-in misfit (syn D2H, adjoint source H2D)."""
    from oracle import ref_qw2d
    if not ref_qw2d.available():
        pytest.skip("compiled reference (oracle/_ref/libqw2d_ref.so) not present")
    from devito_fwi_b200 import configs, fwi
    from devito_fwi_b200.misfit import qWasserstein
    g_true, g_init, g_const, mask = configs.marmousi(nsrc=2, tn=900.)
    obs, dw = fwi.fm_multi(g_true), fwi.fm_multi(g_const)
    nbl = g_init.model.nbl
    x0 = (1. / (g_init.model.vp.data[nbl:-nbl, nbl:-nbl].astype(np.float64) ** 2)).ravel()
    qw = qWasserstein(method='2d', gamma=1.01, num_steps=6, step_scale=4.)
    assert fwi._is_w2d(qw)
    f_dev, g_dev, r_dev = fwi.fwi_loss(x0, g_init, obs, qw, dw, mask, True, True)

    def host_plugin(syn, ob):
        return ref_qw2d.qwasserstein_2d(syn, ob, 1.01, 6, 4.)
    f_ref, g_ref, r_ref = fwi.fwi_loss(x0, g_init, obs, host_plugin, dw, mask, True, True)
    print("fwi_loss QW2D: f %.6e vs %.6e, gradient rel-L2 %.2e, adjoint source rel-L2 %.2e" % (
        f_dev, f_ref, rel_l2(g_dev, g_ref), rel_l2(np.asarray(r_dev[0]), r_ref[0])))
    assert abs(f_dev - f_ref) <= TOL_LOSS * abs(f_ref)
    assert rel_l2(np.asarray(r_dev[0]), r_ref[0]) <= TOL_ADJ and rel_l2(g_dev, g_ref) <= TOL_ADJ
