"""oracle/ref_qw2d.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The 2-D quadratic-Wasserstein misfit of the reference, evaluated by the REFERENCE'S OWN C solver: oracle/_ref/
libqw2d_ref.so is built by oracle/qw2d/Makefile from /root/reference/misfit/QW2D/src/{fot2d.c,normalize.c}, unmodified,
against a stand-in for the absent libfftw3f (oracle/qw2d/fftw3.h, dct_shim.c). Around it, the numpy glue of the
reference restated:
  misfit/misfit.py:18-45   qWasserstein._transform, trans_type='linear'
  misfit/misfit.py:69-79   _2d_calculator (mass, grad / mass)
  misfit/misfit.py:81-104  __call__ (grad * d)
  misfit/bfm.py:155-193    bfmx.setup / solve: float32 records, argv order `n2 n1` = (f.shape[1], f.shape[0]);
                           w2.c:8-75 is replaced by qw2d_ref_gradient() of dct_shim.c (same calls, no files)
Differences from running the reference through its subprocess: the loss is not rounded through "%e" text (w2.c:67),
and fot2d.c runs single-threaded (its OpenMP push-forward is schedule dependent, see oracle/qw2d/Makefile).
Only tests/ and tests/golden/ scripts import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libqw2d_ref.so")
REF_SRC = "/root/reference/misfit/QW2D/src"
_lib = None


def available():
    return os.path.exists(LIB) or os.path.isdir(REF_SRC)


def build():
    """oracle/qw2d/Makefile; needs the reference checkout."""
    subprocess.run(["make", "-C", os.path.join(_HERE, "qw2d")], check=True, stdout=subprocess.DEVNULL)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = ctypes.CDLL(LIB)
        L.qw2d_ref_gradient.restype = ctypes.c_float
        L.qw2d_ref_gradient.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib = L
    return _lib


def bfm_gradient(f, g, num_steps, step_scale):
    """bfmx.gradient (misfit/bfm.py:189-193): (loss, grad) of float32 records f = syn, g = obs of shape (n1_py, n2_py)."""
    f = np.ascontiguousarray(f, dtype=np.float32)
    g = np.ascontiguousarray(g, dtype=np.float32)
    n1_py, n2_py = f.shape
    adj = np.empty_like(f)
    loss = lib().qw2d_ref_gradient(n2_py, n1_py, int(num_steps), ctypes.c_float(step_scale), f.ctypes.data,
                                   g.ctypes.data, adj.ctypes.data)
    return float(loss), adj


def qwasserstein_2d(f, g, gamma=1.0, num_steps=10, step_scale=1.):
    """qWasserstein(trans_type='linear', gamma, method='2d', num_steps, step_scale)(f, g)   [misfit/misfit.py:81-104]."""
    shape = f.shape
    if len(shape) == 1 or shape[1] <= 1:
        raise ValueError("Can not use 2d method for 1D input.")
    min_value = min(f.min(), g.min())
    c = -min_value if min_value < 0 else 0
    c = c * gamma
    mu, nu = f + c, g + c
    d = np.ones(f.shape)
    mass = mu.sum() / mu.size
    loss, grad = bfm_gradient(mu, nu, num_steps, step_scale)
    return loss, (grad / mass) * d
