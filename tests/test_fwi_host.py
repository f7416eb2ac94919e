"""Host-side pieces of the fwi.py mirror (no GPU): the wavelet Filter against golden vectors of the reference's
filter.py, the residual wrapper, and which misfits are moved onto the device."""
import os
import warnings

import numpy as np
import pytest

from devito_fwi_b200 import fwi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "filter_small.npz")

CASES = {
    "bandpass": dict(filter_type="bandpass", freqmin=3.0, freqmax=12.0, corners=10, zerophase=False),
    "bandpass_zerophase": dict(filter_type="bandpass", freqmin=3.0, freqmax=12.0, corners=6, zerophase=True),
    "bandpass_to_highpass": dict(filter_type="bandpass", freqmin=5.0, freqmax=None, corners=4, zerophase=False),
    "lowpass": dict(filter_type="lowpass", freqmax=8.0, corners=10, zerophase=False),
    "highpass": dict(filter_type="highpass", freqmin=4.0, corners=10, zerophase=True),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_filter_matches_reference_golden(name):
    """fwi.Filter == seismic/filter/filter.py (fwi.py:10-44), including the Nyquist fall-back bandpass -> highpass
    (filter.py:55-61) where scipy alone would raise. Fixture: tests/golden/make_filter_golden.py."""
    g = np.load(GOLD)
    df = float(g["df"])
    kw = dict(CASES[name])
    if name == "bandpass_to_highpass":
        kw["freqmax"] = 0.5 * df
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y = fwi.Filter(df=df, **kw)(g["data"])
        y2 = fwi.Filter(df=df, axis=0, **kw)(g["data2"])
    assert np.allclose(y, g[name], rtol=1e-10, atol=1e-12)
    if name != "bandpass_to_highpass":          # the reference's fall-back drops `axis` (filter.py:60-61): 1-D use only
        assert np.allclose(y2, g[name + "_2d_axis0"], rtol=1e-10, atol=1e-12)


def test_filter_errors_like_the_reference():
    with pytest.raises(ValueError):
        fwi.Filter("bandpass", freqmin=3.0, df=100.0)(np.zeros(16))           # missing corner (fwi.py:14-18)
    with pytest.raises(ValueError):
        fwi.Filter("bandpass", freqmin=80.0, freqmax=30.0, df=100.0)(np.zeros(16))   # low corner above Nyquist
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(ValueError):        # lowpass corner clamped to Nyquist (filter.py:136-139): scipy then refuses Wn = 1
            fwi.Filter("lowpass", freqmax=80.0, df=100.0)(np.zeros(16))


def test_lazy_residual_behaves_like_an_ndarray():
    import torch
    t = torch.arange(12, dtype=torch.float32).reshape(4, 3)
    r = fwi.LazyResidual(t)
    ref = t.numpy().copy()
    assert r.shape == (4, 3) and r.dtype == np.float32 and len(r) == 4 and r.size == 12
    assert np.array_equal(np.asarray(r), ref)
    assert np.array_equal(r + 1, ref + 1) and np.array_equal(2 * r, 2 * ref) and np.array_equal(-r, -ref)
    assert np.array_equal(r - ref, np.zeros_like(ref)) and np.array_equal(r.T, ref.T)
    c = r.copy()
    c[0, 0] = 99
    assert r[0, 0] == 0 and np.array_equal(r.astype(np.float64), ref.astype(np.float64))
    assert float(r.sum()) == float(ref.sum()) and np.array_equal(r.ravel(), ref.ravel())
    assert np.linalg.norm(r) == np.linalg.norm(ref)


def test_device_misfits_are_recognised_by_origin_not_by_name():
    def least_square(x, y):                       # a user's own callable that merely shares the name
        return 0.0, x - y
    least_square.__module__ = "user_code"
    assert fwi._is_l2(fwi.least_square) and not fwi._is_l2(least_square)

    class qWasserstein(object):                   # same for a look-alike class defined elsewhere
        method, trans_type, gamma = '1d', 'linear', 1.0
    qWasserstein.__module__ = "user_code"
    assert not fwi._is_w1d(qWasserstein())
    assert not fwi._is_l2(lambda x, y: fwi.least_square(x, y))
    fwi.least_square.b2fwi_host = True            # explicit opt-out
    try:
        assert not fwi._is_l2(fwi.least_square)
    finally:
        del fwi.least_square.b2fwi_host
