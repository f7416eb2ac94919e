"""ctypes binding of libb2fwi.so (the C ABI declared in include/b2fwi.h).

The library is the only compute path of this package: if it is missing the import of
any solver entry point fails loudly -- there is no CPU or PyTorch fallback.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libb2fwi.so")
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["api.cu", "stream_kernels.cu", "stream_tma.cu", "resident2d.cu", "resident2d_lat.cu", "resident2d_lat_r2.cu",
           "resident2d_lat_r3.cu", "resident2d_lat_r4.cu", "res2d_api.cu", "qw2d.cu"]
# per-source extra flags: the QW2D solver follows the reference's C arithmetic, which is built without FMA contraction
EXTRA_FLAGS = {"qw2d.cu": ["-fmad=false"]}
LINK_FLAGS = ["-lcufft", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-I" + os.path.join(_ROOT, "include")]


class B2fwiError(RuntimeError):
    pass


class Grid(ctypes.Structure):
    """struct b2fwi_grid"""
    _fields_ = [("ndim", ctypes.c_int32), ("shape", ctypes.c_int32 * 3),
                ("space_order", ctypes.c_int32), ("halo", ctypes.c_int32),
                ("spacing", ctypes.c_float * 3), ("origin", ctypes.c_float * 3), ("fs", ctypes.c_int32),
                ("kernel", ctypes.c_int32)]


class Sparse(ctypes.Structure):
    """struct b2fwi_sparse (device pointers)"""
    _fields_ = [("npoint", ctypes.c_int32), ("ncorner", ctypes.c_int32),
                ("corner_off", ctypes.c_void_p), ("corner_w", ctypes.c_void_p),
                ("ncell", ctypes.c_int32),
                ("cell_off", ctypes.c_void_p), ("cell_ptr", ctypes.c_void_p),
                ("contrib_pt", ctypes.c_void_p), ("contrib_w", ctypes.c_void_p),
                ("row_tile", ctypes.c_int32),
                ("con_rowptr", ctypes.c_void_p), ("con_off", ctypes.c_void_p), ("max_row_con", ctypes.c_int32),
                ("pt_order", ctypes.c_void_p), ("pt_home", ctypes.c_void_p), ("pt_rowptr", ctypes.c_void_p),
                ("z_min", ctypes.c_int32), ("z_max", ctypes.c_int32),
                ("r_min", ctypes.c_int32), ("r_max", ctypes.c_int32), ("p_min", ctypes.c_int32), ("p_max", ctypes.c_int32)]


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(_ROOT, "include", "b2fwi.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    """Compile csrc/*.cu for sm_100a into devito_fwi_b200/libb2fwi.so (in-tree): one nvcc -c per source in
    parallel, then one link step."""
    if not force and not needs_build():
        return LIB_PATH
    import tempfile
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    if verbose:
        flags = ["-Xptxas=-v"] + flags
    with tempfile.TemporaryDirectory() as tmp:
        procs = []
        for src in sources():
            obj = os.path.join(tmp, os.path.basename(src)[:-3] + ".o")
            cmd = ["nvcc"] + flags + EXTRA_FLAGS.get(os.path.basename(src), []) + ["-c", "-o", obj, src]
            procs.append((obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs, log = [], ""
        for obj, pr in procs:
            out, _ = pr.communicate()
            log += out
            if pr.returncode != 0:
                raise B2fwiError("nvcc failed:\n" + log)
            objs.append(obj)
        res = subprocess.run(["nvcc"] + NVCC_FLAGS + ["-o", LIB_PATH] + objs + LINK_FLAGS, stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            raise B2fwiError("nvcc link failed:\n" + res.stdout)
        if verbose:
            print(log + res.stdout)
    return LIB_PATH


_lib = None

_P = ctypes.c_void_p
_I = ctypes.c_int32
_F = ctypes.c_float
_G = ctypes.POINTER(Grid)
_S = ctypes.POINTER(Sparse)

PROTOTYPES = {
    # name: (restype, argtypes)   -- must list every symbol declared in include/b2fwi.h
    "b2fwi_version": (_I, []),
    "b2fwi_last_error": (ctypes.c_char_p, []),
    "b2fwi_launch_count": (ctypes.c_int64, []),
    "b2fwi_set_option": (ctypes.c_int, [ctypes.c_char_p, _I]),
    "b2fwi_field_layout": (ctypes.c_int, [_G, ctypes.POINTER(ctypes.c_int64 * 3),
                                          ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    "b2fwi_prepare_coeffs": (ctypes.c_int, [_G, _P, _P, _F, _P, _P]),
    "b2fwi_forward": (ctypes.c_int, [_G, _P, _P, _F, _I, _I, _I, _P, _S, _P, _S, _P, _I, _P, _P, _I, _P]),
    "b2fwi_gradient": (ctypes.c_int, [_G, _P, _P, _F, _I, _I, _I, _P, _S, _P, _I, _I, _P, _P, _P]),
    "b2fwi_adjoint": (ctypes.c_int, [_G, _P, _P, _F, _I, _I, _I, _P, _S, _P, _S, _P, _P]),
    "b2fwi_born": (ctypes.c_int, [_G, _P, _P, _F, _I, _I, _I, _P, _S, _P, _S, _P, _P, _P, _P, _P]),
    "b2fwi_geometry_mask": (ctypes.c_int, [_G, _I, _P, _I, _P, _P]),
    "b2fwi_crop_mask_accumulate": (ctypes.c_int, [_G, _I, _P, _P, _P, _P]),
    "b2fwi_res2d_plan_model": (ctypes.c_int, [_G, _I, _I, _I, _P]),
    "b2fwi_res2d_plan_exact": (ctypes.c_int, [_G, _I, _I, _I, _P]),
    "b2fwi_res2d_max_active_clusters": (ctypes.c_int, [_G, _P, _P]),
    "b2fwi_res2d_prepare": (ctypes.c_int, [_G, _P, _F, _P, _P]),
    "b2fwi_res2d_forward": (ctypes.c_int, [_G, _P, _P, _P, _P, _F, _I, _I, _I, _I, _P, _I, _P, _P, _I, _P, _P, _P]),
    "b2fwi_res2d_gradient": (ctypes.c_int, [_G, _P, _P, _P, _P, _F, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P]),
    "b2fwi_window_mask_accumulate": (ctypes.c_int, [_I, _I, _P, ctypes.c_int64, _I, _P, _P, _P]),
    "b2fwi_window_mask_accumulate_batch": (ctypes.c_int, [_I, _I, _I, _P, ctypes.c_int64, ctypes.c_int64, _I, _P, _P, _P]),
    "b2fwi_w1d_misfit": (ctypes.c_int, [_P, _P, _P, _I, _I, _I, ctypes.c_double, _P, _P, _P, _P]),
    "b2fwi_w1d_scratch_bytes": (ctypes.c_int64, [_I, _I, _I]),
    "b2fwi_l2_misfit": (ctypes.c_int, [_P, _P, _P, ctypes.c_int64, _P, _P, _P, _P]),
    "b2fwi_qw2d_misfit": (ctypes.c_int, [_P, _P, _P, _I, _I, _I, ctypes.c_double, _I, _F, _P, _P, _P, _P, _P]),
    "b2fwi_qw2d_scratch_bytes": (ctypes.c_int64, [_I, _I, _I]),
    "b2fwi_qw2d_debug_step": (ctypes.c_int, [_I, _I, _I, _P, _P, _P, _P, _F, _P, _P, _P, _P]),
}


def lib():
    """Load libb2fwi.so; raises B2fwiError when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2fwiError(
                "libb2fwi.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'`; this package has no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().b2fwi_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError("b2fwi: " + msg)
        raise B2fwiError("b2fwi error %d: %s" % (rc, msg))
