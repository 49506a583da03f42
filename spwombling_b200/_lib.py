"""ctypes binding of libspwombling.so (the C ABI declared in include/spwombling.h).

There is no CPU fallback: if the shared library is missing, or an entry point reports an error, the
caller gets an exception.  The library is built in-tree by ``spwombling_b200.build``.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libspwombling.so")

SPW_OK, SPW_ERR_NOT_PD, SPW_ERR_ARG, SPW_ERR_CUDA = 0, 1, 2, 3
COV_CODES = {"gaussian": 0, "matern1": 1, "matern2": 2}

c_int, c_ll, c_dbl, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_double, ctypes.c_void_p
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)

# name -> (restype, argtypes); must list every symbol of include/spwombling.h (tests/test_abi.py checks)
SIGNATURES = {
    "spw_version": (c_int, []),
    "spw_last_error": (ctypes.c_char_p, []),
    "spw_device_count": (c_int, []),
    "spw_set_device": (c_int, [c_int]),
    "spw_shutdown": (c_int, []),
    "spw_set_workspace_limit": (c_int, [c_ll]),
    "spw_launch_count": (c_ll, [c_int]),
    "spw_last_gather_used_nccl": (c_int, []),
    "spw_profile_enable": (c_int, [c_int]),
    "spw_profile_read": (c_int, [c_vp, c_vp, c_vp, c_int]),
    "spw_debug_diag_timing": (c_int, [c_vp]),
    "spw_debug_dag_trace": (c_ll, [c_int, c_int, c_vp, c_ll]),
    "spw_debug_potrf_time": (c_int, [c_int, c_int, c_int, c_vp, c_vp]),
    "spw_debug_la_trace": (c_ll, [c_int, c_int, c_vp, c_ll]),
    "spw_cov": (c_int, [c_int, c_int, c_vp, c_dbl, c_dbl, c_vp]),
    "spw_cov_delta": (c_int, [c_int, c_ll, c_vp, c_dbl, c_dbl, c_vp]),
    "spw_cross_cov": (c_int, [c_int, c_int, c_vp, c_int, c_vp, c_dbl, c_vp]),
    "spw_chol": (c_int, [c_int, c_vp, c_vp, c_vp, c_vp]),
    "spw_hlm_run": (c_int, [c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl,
                            c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "spw_gradient_run": (c_int, [c_int, c_int, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_dbl,
                                 c_vp, c_vp, c_vp, c_vp, c_vp]),
    "spw_gradient_shard_dev": (c_int, [c_int, c_int, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp,
                                       c_vp, c_vp, c_vp, c_vp]),
    "spw_spsample_run": (c_int, [c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "spw_wombling_rect_run": (c_int, [c_int, c_int, c_vp, c_int, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                      c_vp, c_vp]),
    "spw_hpd_summary": (c_int, [c_int, c_int, c_int, c_dbl, c_vp, c_vp]),
    "spw_hpd_summary_dev": (c_int, [c_int, c_int, c_int, c_int, c_dbl, c_vp, c_vp, c_vp]),
    "spw_draws_to_r_layout_dev": (c_int, [c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "spw_surface_analysis": (c_int, [c_int, c_int, c_vp, c_int, c_dbl, c_vp, c_vp]),
    "spw_surface_analysis_dev": (c_int, [c_int, c_int, c_vp, c_vp, c_vp]),
    "spw_cov_p": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "spw_hpd_summary_p": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "spw_hlm_run_p": (c_int, [c_vp] * 19),
    "spw_gradient_run_p": (c_int, [c_vp] * 18),
}

_lib = None


class SpwError(RuntimeError):
    def __init__(self, code, message, info=None):
        super().__init__(message)
        self.code = code
        self.info = info


class NotPositiveDefinite(SpwError):
    """chol() failed -- the reference raises R's 'leading minor ... is not positive' error here."""


def load():
    """Load the shared library (once).  Raises if it has not been built: no fallback path exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # not built yet: build it in-tree if nvcc is here (this is a compile step, not a fallback path)
        import shutil
        if shutil.which(os.environ.get("NVCC", "nvcc")):
            from .build import build
            build()
    if not os.path.exists(LIB_PATH):
        raise ImportError("libspwombling.so not found at %s -- run `python -m spwombling_b200.build` "
                          "(nvcc, sm_100a); there is no CPU fallback" % LIB_PATH)
    if "SPW_NCCL_PATH" not in os.environ:
        # the in-process multi-GPU gather dlopen()s NCCL at run time: prefer the copy PyTorch bundles (found without
        # importing torch), or a later `import torch` would meet the older system libnccl.so.2 under the same soname
        try:
            import importlib.util
            spec = importlib.util.find_spec("nvidia.nccl")
            for loc in (spec.submodule_search_locations or []) if spec else []:
                cand = os.path.join(loc, "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    os.environ["SPW_NCCL_PATH"] = cand
                    break
        except Exception:       # pragma: no cover
            pass
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().spw_last_error().decode("utf-8", "replace")


def check(rc, info=None):
    if rc == SPW_OK:
        return
    msg = last_error()
    if rc == SPW_ERR_NOT_PD:
        raise NotPositiveDefinite(rc, msg, None if info is None else [int(v) for v in info])
    if rc == SPW_ERR_ARG:
        raise ValueError(msg)
    raise SpwError(rc, "libspwombling: " + msg)


def ptr(a):
    """Host pointer of a C- or F-contiguous float64 / int32 numpy array (None -> NULL)."""
    if a is None:
        return None
    return a.ctypes.data_as(c_vp)


def f64(a, order="F"):
    return np.require(a, dtype=np.float64, requirements=["F" if order == "F" else "C", "A"])
