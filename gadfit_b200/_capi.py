"""ctypes binding of the C-ABI declared in ``include/gadfit_b200.h``.

The binding is generic over (shared library, symbol prefix): the product
library exports ``gadfit_b200_*``; the test-only oracle libraries export the
same call shapes under other prefixes and are bound by ``tests/`` through the
same class, never by this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GADFIT_B200_LIB: developer override to A/B two builds of the same library
LIB_PATH = os.environ.get("GADFIT_B200_LIB") or os.path.join(_HERE, "libgadfit_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class Settings(C.Structure):
    """POD mirror of ``gadfit_b200_settings`` (reference lm_solver.h:198-210)."""

    _fields_ = [
        ("iteration_limit", C.c_int),
        ("lambda_incs", C.c_int),
        ("lambda_down", C.c_double),
        ("lambda_up", C.c_double),
        ("damp_max", C.c_int),
        ("DTD_min", c_double_p),
        ("n_DTD_min", C.c_int),
        ("acceleration_threshold", C.c_double),
        ("verbosity", C.c_uint),
        ("loss", C.c_int),
        ("n_threads", C.c_int),
    ]


class Trial(C.Structure):
    _fields_ = [
        ("lambda_", C.c_double),
        ("new_chi2", C.c_double),
        ("old_chi2", C.c_double),
        ("iteration", C.c_int),
        ("decision", C.c_int),
    ]


# gadfit_b200_log_fn
LOG_FN = C.CFUNCTYPE(None, C.c_char_p, C.c_void_p)


class GadfOptions(C.Structure):
    """gadfit_b200_gadf_options: the Fortran-only arguments of gadf_fit"""
    _fields_ = [
        ("enabled", C.c_int),
        ("uphill", C.c_int),
        ("nielsen", C.c_int),
        ("umnigh", C.c_int),
        ("chi2_abs", C.c_double),
        ("chi2_rel", C.c_double),
        ("grad_chi2", C.c_double),
        ("cos_phi", C.c_double),
        ("rel_error", C.c_double),
        ("rel_error_global", C.c_double),
    ]


class Timings(C.Structure):
    _fields_ = [("seconds", C.c_double * 5), ("calls", C.c_int * 5)]


# status codes of include/gadfit_b200.h
OK = 0
UNKNOWN_OPERATION = 1
LATE_ADD_DATASET_CALL = 2
SET_PAR_INVALID_INDEX = 3
UNINITIALIZED_PARAMETER = 4
NEGATIVE_DEGREES_OF_FREEDOM = 5
NO_FITTING_PARAMETERS = 6
NO_GLOBAL_PARAMETERS = 7
INSUFFICIENT_INTEGRATION_WORKSPACE = 8
UNKNOWN_MODEL = 9
CUDA_ERROR = 10
NCCL_ERROR = 11
INVALID_ARGUMENT = 12
NOT_AVAILABLE = 13
TOO_MANY_PARAMETERS = 14

# gadfit_b200_array
JACOBIAN, JTJ, DTD, LEFT_SIDE, RIGHT_SIDE, RESIDUALS, INV_JTJ, DELTA1, DELTA2, OMEGA, INV_JTJ_DIAGONAL = range(11)

# every symbol include/gadfit_b200.h declares (checked by tests/test_capi_host.py)
CORE_SYMBOLS = [
    "create", "destroy", "add_dataset", "set_par", "default_settings",
    "set_settings", "fit", "get_par", "get_value", "chi2", "dof", "get_array",
    "get_trace", "n_iterations", "get_timings", "last_error",
]
PRODUCT_ONLY_SYMBOLS = [
    "default_gadf_options", "set_gadf_options", "get_stop_reason", "set_log",
    "n_models", "model_info", "load_models", "n_launches", "jacobian_kernel_time",
    "status_name", "set_device", "upload", "invalidate_data", "set_phase_timers", "comm_unique_id", "set_comm",
    "partition", "measure_fp64_peak", "microbench", "comm_stats",
]


class Api:
    """Typed function table of one library/prefix pair."""

    def __init__(self, path: str, prefix: str, product: bool):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing: build it with `python -c 'import "
                f"__graft_entry__ as g; g.build()'` (there is no CPU fallback)")
        # RTLD_LOCAL: the oracle libraries export C++ symbols of the reference
        # that must not interpose each other
        self.lib = C.CDLL(path, mode=getattr(os, "RTLD_NOW", 2) | getattr(os, "RTLD_LOCAL", 0))
        self.path = path
        self.prefix = prefix
        self.product = product
        vp = C.c_void_p
        sig = {
            "create": ([C.c_char_p, C.POINTER(vp)], C.c_int),
            "destroy": ([vp], C.c_int),
            "add_dataset": ([vp, C.c_int, vp, vp, vp], C.c_int),
            "set_par": ([vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_char_p], C.c_int),
            "default_settings": ([C.POINTER(Settings)], C.c_int),
            "set_settings": ([vp, C.POINTER(Settings)], C.c_int),
            "fit": ([vp, C.c_double], C.c_int),
            "get_par": ([vp, C.c_int, C.c_int, c_double_p], C.c_int),
            "get_value": ([vp, C.c_double, C.c_int, c_double_p], C.c_int),
            "chi2": ([vp, c_double_p], C.c_int),
            "dof": ([vp, c_int_p], C.c_int),
            "get_array": ([vp, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)], C.c_int),
            "get_trace": ([vp, C.POINTER(Trial), C.c_size_t, C.POINTER(C.c_size_t)], C.c_int),
            "n_iterations": ([vp, c_int_p], C.c_int),
            "get_timings": ([vp, C.POINTER(Timings)], C.c_int),
            "last_error": ([vp], C.c_char_p),
        }
        self.has_ref_log = prefix == "gadfit_ref_"
        if self.has_ref_log:
            sig["get_log"] = ([vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)], C.c_int)
        self.has_gadf = product or prefix == "gadfit_oracle_"
        if self.has_gadf:
            sig.update({
                "default_gadf_options": ([C.POINTER(GadfOptions)], C.c_int),
                "set_gadf_options": ([vp, C.POINTER(GadfOptions)], C.c_int),
                "get_stop_reason": ([vp, c_int_p], C.c_int),
            })
        if product:
            sig.update({
                "set_log": ([vp, LOG_FN, vp], C.c_int),
                "n_models": ([], C.c_int),
                "model_info": ([C.c_int, C.POINTER(C.c_char_p), c_int_p], C.c_int),
                "load_models": ([C.c_char_p], C.c_int),
                "n_launches": ([vp, C.POINTER(C.c_longlong)], C.c_int),
                "jacobian_kernel_time": ([vp, c_double_p, c_int_p], C.c_int),
                "status_name": ([C.c_int], C.c_char_p),
                "set_device": ([vp, C.c_int], C.c_int),
                "upload": ([vp], C.c_int),
                "invalidate_data": ([vp], C.c_int),
                "set_phase_timers": ([vp, C.c_int], C.c_int),
                "comm_unique_id": ([vp], C.c_int),
                "set_comm": ([vp, C.c_int, C.c_int, vp], C.c_int),
                "partition": ([C.c_int, c_int_p, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p], C.c_int),
                "measure_fp64_peak": ([c_double_p], C.c_int),
                "microbench": ([C.c_int, c_double_p], C.c_int),
                "comm_stats": ([vp, c_int_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)], C.c_int),
            })
        for name, (argtypes, restype) in sig.items():
            fn = getattr(self.lib, prefix + name)
            fn.argtypes = argtypes
            fn.restype = restype
            setattr(self, name, fn)


_product_api: Api | None = None


def product_api() -> Api:
    """The product library; fails loudly when the CUDA extension is missing."""
    global _product_api
    if _product_api is None:
        _product_api = Api(LIB_PATH, "gadfit_b200_", product=True)
    return _product_api
