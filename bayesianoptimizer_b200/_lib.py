"""ctypes binding of ``libbo_b200.so`` (C ABI declared in ``include/bo_b200.h``).

There is deliberately no fallback: if the CUDA library is missing the import of the product path
fails loudly (``BoLibraryError``), and every compute entry point returns BO_E_CUDA without a B200.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BO_B200_LIB: A/B builds of the same ABI for the profiling tools (csrc/Makefile `alt`); the product loads libbo_b200.so
LIB_PATH = os.environ.get("BO_B200_LIB") or os.path.join(_HERE, "libbo_b200.so")

BO_MAX_DIM = 16
BO_MAX_TOPK = 64
BO_MAX_SELECT = 8192
BO_SOBOL_BITS = 30

KERNEL_MATERN52, KERNEL_RBF, KERNEL_LINEAR_MATERN52 = 0, 1, 2
ACQ_EI, ACQ_LOGEI, ACQ_UCB, ACQ_VAR, ACQ_MEAN = 0, 1, 2, 3, 4
E_INVALID, E_CUDA, E_NOMEM, E_NOTFIT, E_CAPACITY = -1, -2, -3, -4, -5


class BoLibraryError(RuntimeError):
    pass


class BoSobol(C.Structure):
    _fields_ = [("d", C.c_int32),
                ("direction", (C.c_uint32 * BO_SOBOL_BITS) * BO_MAX_DIM),
                ("shift", C.c_uint32 * BO_MAX_DIM)]


_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_pd = C.POINTER(C.c_double)

# name -> (restype, argtypes); every symbol include/bo_b200.h declares
SIGNATURES = {
    "bo_abi_version": (C.c_int, []),
    "bo_device_count": (C.c_int, []),
    "bo_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "bo_destroy": (None, [_vp]),
    "bo_last_error": (C.c_char_p, [_vp]),
    "bo_release_workspace": (C.c_int, [_vp]),
    "bo_fit": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _pd, _f64, _f64, _f64, _f64, _vp]),
    "bo_fit_host": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _pd, _f64, _f64, _f64, _f64, _vp]),
    "bo_fit_ex": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _pd, _f64, _f64, _f64, _f64, _f64, _i32, _vp]),
    "bo_posterior_multi": (C.c_int, [_vp, _vp, _i32, _pd, _vp, _i64, _f64, _vp, _vp, _vp]),
    "bo_topk_scores": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _vp, _vp, _vp]),
    "bo_svgp_load": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _pd, _f64, _f64, _f64, _f64, _f64, _vp, _vp, _vp]),
    "bo_get_state": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "bo_num_obs": (C.c_int, [_vp]),
    "bo_posterior": (C.c_int, [_vp, _vp, _i64, _f64, _vp, _vp, _vp]),
    "bo_sweep": (C.c_int, [_vp, _i32, _f64, _f64, _f64, _vp, C.POINTER(BoSobol), _i64, _i64, _i32,
                           _vp, _vp, _vp, _vp, _vp, _vp]),
    "bo_sweep_host": (C.c_int, [_vp, _i32, _f64, _f64, _f64, _vp, C.POINTER(BoSobol), _i64, _i64, _i32,
                                _vp, _vp, _vp]),
    "bo_sobol_points": (C.c_int, [_vp, C.POINTER(BoSobol), _vp, _i64, _vp, _vp]),
    "bo_refine": (C.c_int, [_vp, _i32, _f64, _f64, _f64, _vp, _i32, _i32, _vp, _vp, _vp]),
    "bo_acq_grad": (C.c_int, [_vp, _i32, _f64, _f64, _f64, _vp, _i32, _vp, _vp, _vp]),
    "bo_append": (C.c_int, [_vp, _vp, _f64, _i32, _vp]),
    "bo_lml_grad_batched": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _f64, _vp, _i32, _vp, _vp, _vp, _vp]),
    "bo_fps": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i64, _vp, _vp]),
    "bo_fp64_peak": (C.c_int, [_vp, _i32, _f64, _pd]),
    "bo_gemm_probe": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _pd]),
    "bo_set_sweep_mode": (C.c_int, [_vp, _i32]),
    "bo_resolve_sweep_mode": (C.c_int, [_vp, _i64]),
    "bo_set_linear_variance_ard": (C.c_int, [_vp, _pd, _i32]),
    "bo_last_sweep_path": (C.c_int, [_vp]),
    "bo_last_sweep_flagged": (C.c_int64, [_vp]),
    "bo_i8_peak": (C.c_int, [_vp, _f64, _pd]),
    "bo_launch_count": (C.c_int64, [_vp]),
    "bo_last_sweep_ms": (C.c_double, [_vp]),
}

_lib = None


def load():
    """Load the shared library once; raise BoLibraryError (never fall back) if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BoLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). The B200 GP path has no CPU fallback.")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise BoLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise BoLibraryError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.bo_abi_version() != 1:
        raise BoLibraryError("ABI version mismatch between bo_b200.h and libbo_b200.so")
    _lib = lib
    return lib
