#!/usr/bin/env python
"""Golden (f, g) vectors of the reference's OWN fwi.py glue code.

The reference's `fwi.py` (fwi_obj_single / fwi_obj_multi / fix_source_illumination, fwi.py:104-205) is imported
UNMODIFIED from /root/reference and run with
  * the reference's own `misfit.least_square` (misfit/misfit.py:5-9),
  * host objects (Model, Receiver, AcquisitionGeometry, Function) from this repo's numpy-only mirror,
  * and, in place of Devito's propagator, the pinned CPU oracle (fp64) behind the AcousticWaveSolver interface.
So the crop, the axis-swapped source/receiver muting, the illumination sum over the saved wavefield, the
preconditioning, the mask and the shot sum are the REFERENCE's code; only the wave propagation is the oracle
(which tests/test_oracle_kat.py pins to the reference's known-answer values).

Run in the build container (needs /root/reference):   python tests/golden/make_fwi_golden.py
Writes tests/golden/fwi_obj_small.npz; tests/test_gpu_golden.py compares the CUDA path against it.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

import devito_fwi_b200 as b                      # noqa: E402  (host-side mirror: numpy only on CPU)
from oracle import ref                           # noqa: E402
from tests.util import ref_model                 # noqa: E402


class OracleSolver(object):
    """AcousticWaveSolver look-alike on the CPU oracle (fp64), for the reference's fwi.py call sites
    (fwi.py:134-137,162-163)."""

    def __init__(self, model, geometry, space_order=4, **kwargs):
        self.model, self.geometry, self.space_order = model, geometry, space_order

    def _rm(self, vp):
        rm = ref_model(self.model, np.float64)
        rm.vp = np.array(vp.data, dtype=np.float64)
        return rm

    def forward(self, vp=None, save=None, **kwargs):
        g = self.geometry
        rec = g.rec
        d, u = ref.forward(self._rm(vp or self.model.vp), g.src_positions, g.rec_positions,
                           np.float64(g.src.data), g.nt, float(g.dt), save=bool(save), space_order=self.space_order)
        rec.data[:] = d
        wfd = types.SimpleNamespace(data=u)
        return rec, wfd, None

    def gradient(self, rec, u, vp=None, grad=None, **kwargs):
        g = self.geometry
        out = ref.gradient(self._rm(vp or self.model.vp), np.float64(rec.data), g.rec_positions, u.data, g.nt,
                           float(g.dt), space_order=self.space_order)
        grad.data[:] = grad.data + out
        return grad, None

    jacobian_adjoint = gradient


def install_shims():
    """Module names the reference's fwi.py imports (fwi.py:1-8), bound to the objects above."""
    devito = types.ModuleType("devito")
    devito.Function = b.Function
    seismic = types.ModuleType("seismic")
    seismic.Model, seismic.Receiver, seismic.AcquisitionGeometry = b.Model, b.Receiver, b.AcquisitionGeometry
    acoustic = types.ModuleType("seismic.acoustic")
    acoustic.AcousticWaveSolver = OracleSolver
    flt = types.ModuleType("seismic.filter")
    flt.bandpass = flt.lowpass = flt.highpass = None      # only used with --filter 1
    distributed = types.ModuleType("distributed")
    distributed.wait = lambda futures: None               # dead dask path (fwi.py:83-102)
    w2 = types.ModuleType("w2")
    w2.BFM = object                                       # misfit/bfm.py:1 (module absent from the reference tree)
    seismic.acoustic, seismic.filter = acoustic, flt
    sys.modules.update({"devito": devito, "seismic": seismic, "seismic.acoustic": acoustic,
                        "seismic.filter": flt, "distributed": distributed, "w2": w2})


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def problem():
    """Small survey with off-grid sources / receivers (so that the bilinear weights and the mute matter)."""
    shape, nbl, so = (61, 41), 10, 4
    rng = np.random.default_rng(11)
    xx, zz = np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), indexing="ij")
    vp_true = (1.5 + 0.03 * zz + 0.4 * np.exp(-((xx - 30) ** 2 + (zz - 22) ** 2) / 40.)).astype(np.float32)
    vp_init = (1.5 + 0.03 * zz).astype(np.float32)
    src = np.stack([np.array([83.3, 301.7, 512.9]), np.full(3, 23.1)], axis=1)
    rec = np.stack([np.linspace(14.2, 588.8, 13), np.full(13, 31.7)], axis=1)
    mask = np.ones(shape, dtype=np.float32)
    mask[:, :4] = 0
    kw = dict(origin=(0., 0.), spacing=(10., 10.), shape=shape, space_order=so, nbl=nbl, dt=1.5)
    return dict(vp_true=vp_true, vp_init=vp_init, vp_const=np.full(shape, 1.5, np.float32), src=src, rec=rec,
                mask=mask, kw=kw, t0=0., tn=330., f0=0.02)


def geometries(p):
    out = []
    for key in ("vp_true", "vp_init", "vp_const"):
        model = b.Model(vp=p[key], **p["kw"])
        out.append(b.AcquisitionGeometry(model, p["rec"], p["src"], p["t0"], p["tn"], f0=p["f0"], src_type="Ricker"))
    return out


def main():
    install_shims()
    ref_fwi = load("reference_fwi", os.path.join(REF, "fwi.py"))
    ref_misfit = load("reference_misfit", os.path.join(REF, "misfit", "misfit.py")) \
        if False else None
    # misfit/misfit.py does a relative import of .bfm: import it as a package instead
    sys.path.insert(0, REF)
    import misfit as reference_misfit_pkg
    p = problem()
    g_true, g_init, g_const = geometries(p)
    obs = ref_fwi.fm_multi(g_true)
    dw = ref_fwi.fm_multi(g_const)
    out = {}
    for tag, direct, mask, precond in (("full", dw, p["mask"], True), ("plain", None, None, False)):
        f, g, res = ref_fwi.fwi_obj_multi(g_init, obs, reference_misfit_pkg.least_square, direct, mask, precond, True)
        out["f_" + tag], out["g_" + tag] = np.float64(f), np.asarray(g, dtype=np.float64)
        out["res0_" + tag] = np.asarray(res[0], dtype=np.float32)
    x = (1. / (np.float64(p["vp_init"]) ** 2)).ravel()
    f, g, _ = ref_fwi.fwi_loss(x, g_init, obs, reference_misfit_pkg.least_square, dw, p["mask"], True, True)
    out["f_loss"], out["g_loss"] = np.float64(f), np.asarray(g, dtype=np.float64)
    out["obs0"] = np.asarray(obs[0].data, dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "fwi_obj_small.npz"), **out)
    print({k: (v.shape if getattr(v, "shape", ()) else float(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
