"""Golden vectors for fwi.Filter: the reference's unmodified seismic/filter/filter.py (bandpass / lowpass / highpass and
their Nyquist edge branches, filter.py:33-182) driven through the reference's own seismic_filter dispatch
(fwi.py:10-29, restated here because importing the reference's fwi.py needs devito).
Run in the build container (needs /root/reference):  python tests/golden/make_filter_golden.py"""
import importlib.util
import os
import warnings

import numpy as np

spec = importlib.util.spec_from_file_location("ref_filter", "/root/reference/seismic/filter/filter.py")
rf = importlib.util.module_from_spec(spec)
spec.loader.exec_module(rf)

rng = np.random.default_rng(11)
nt = 257
data = rng.standard_normal(nt).cumsum()
data2 = rng.standard_normal((nt, 3)).cumsum(axis=0)
df = 1000 / 2.95
cases = {
    "bandpass": dict(filter_type="bandpass", freqmin=3.0, freqmax=12.0, corners=10, zerophase=False),
    "bandpass_zerophase": dict(filter_type="bandpass", freqmin=3.0, freqmax=12.0, corners=6, zerophase=True),
    "bandpass_to_highpass": dict(filter_type="bandpass", freqmin=5.0, freqmax=0.5 * df, corners=4, zerophase=False),
    "lowpass": dict(filter_type="lowpass", freqmax=8.0, corners=10, zerophase=False),
    "highpass": dict(filter_type="highpass", freqmin=4.0, corners=10, zerophase=True),
}
out = {"data": data, "data2": data2, "df": df}
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for name, kw in cases.items():
        t = kw["filter_type"]
        for tag, d in (("", data), ("_2d_axis0", data2)):
            axis = -1 if d.ndim == 1 else 0
            if t == "bandpass":
                y = rf.bandpass(d, kw["freqmin"], kw["freqmax"], df, kw["corners"], kw["zerophase"], axis)
            elif t == "lowpass":
                y = rf.lowpass(d, kw["freqmax"], df, kw["corners"], kw["zerophase"], axis)
            else:
                y = rf.highpass(d, kw["freqmin"], df, kw["corners"], kw["zerophase"], axis)
            out[name + tag] = y
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "filter_small.npz"), **out)
print("wrote", sorted(out))
