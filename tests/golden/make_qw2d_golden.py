"""Golden vectors for the 2-D quadratic-Wasserstein misfit: the reference's own solver (misfit/QW2D/src/fot2d.c +
normalize.c compiled unmodified by oracle/qw2d/Makefile against the DCT stand-in for the absent libfftw3f) under the
numpy glue of misfit/misfit.py restated in oracle/ref_qw2d.py.
Run in the build container (needs /root/reference):  python tests/golden/make_qw2d_golden.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_qw2d  # noqa: E402


def records(nt, nrec, seed):
    """Two band-limited 'shot records' (moved-out wavelets + a weak random field), float32, sign-changing."""
    rng = np.random.default_rng(seed)
    t = np.arange(nt, dtype=np.float64)[:, None]
    x = np.arange(nrec, dtype=np.float64)[None, :]

    def rec(shift, v):
        t0 = 0.25 * nt + shift + np.sqrt(1.0 + (v * (x - 0.45 * nrec)) ** 2)
        w = (t - t0) / (0.03 * nt)
        return (1.0 - 2.0 * w ** 2) * np.exp(-w ** 2) + 0.4 * np.exp(-((t - 0.7 * nt - 0.1 * x) / (0.05 * nt)) ** 2)

    noise = 0.02 * rng.standard_normal((nt, nrec))
    f = rec(0.0, 0.9) + noise
    g = rec(0.04 * nt, 1.1) + 0.5 * noise
    return f.astype(np.float32), g.astype(np.float32)


if __name__ == "__main__":
    out = {}
    for name, (nt, nrec, steps, scale, gamma) in {"a": (96, 33, 15, 4.0, 1.01), "b": (61, 48, 6, 1.0, 1.0)}.items():
        f, g = records(nt, nrec, 3)
        loss, adj = ref_qw2d.qwasserstein_2d(f, g, gamma, steps, scale)
        out.update({name + "_f": f, name + "_g": g, name + "_loss": loss, name + "_adj": adj,
                    name + "_par": np.array([steps, scale, gamma])})
        print(name, f.shape, "loss %.8e" % loss, "max|adj| %.4e" % np.abs(adj).max())
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "qw2d_small.npz"), **out)
