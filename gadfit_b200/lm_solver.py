"""Python mirror of the reference solver interface ``gadfit::LMsolver``.

Same method names, argument meaning, defaults and error behaviour as the
reference class (reference c++/gadfit/lm_solver.h:146-211), implemented as a
thin layer over the C-ABI (``include/gadfit_b200.h``).  All arithmetic happens
in the CUDA library; this file only marshals pointers.

    solver = LMsolver("exponential")          # LMsolver solver { exponential };
    solver.addDataset(x1, y1)                  # solver.addDataset(x_data_1, y_data_1);
    solver.addDataset(x2, y2)
    solver.setPar(0, 1.0, True, 0)             # solver.setPar(0, 1.0, true, 0);
    solver.setPar(1, 1.0, True)                # global parameter
    solver.settings.iteration_limit = 4
    solver.fit(1.0)
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import _capi
from ._capi import Api, Settings, Timings, Trial


# Exceptions: same names as reference c++/gadfit/exceptions.h:21-80 ----------

class GADfitException(Exception):
    pass


class UnknownOperation(GADfitException):
    pass


class LateAddDatasetCall(GADfitException):
    pass


class SetParInvalidIndex(GADfitException):
    pass


class UninitializedParameter(GADfitException):
    pass


class NegativeDegreesOfFreedom(GADfitException):
    pass


class NoFittingParameters(GADfitException):
    pass


class NoGlobalParameters(GADfitException):
    pass


class InsufficientIntegrationWorkspace(GADfitException):
    pass


class UnknownModel(GADfitException):
    pass


class CudaError(GADfitException):
    pass


class NcclError(GADfitException):
    pass


class InvalidArgument(GADfitException):
    pass


class TooManyParameters(GADfitException):
    """more than 32 parameters in a fitting function"""


class NotAvailable(GADfitException):
    pass


_STATUS_TO_EXCEPTION = {
    _capi.UNKNOWN_OPERATION: UnknownOperation,
    _capi.LATE_ADD_DATASET_CALL: LateAddDatasetCall,
    _capi.SET_PAR_INVALID_INDEX: SetParInvalidIndex,
    _capi.UNINITIALIZED_PARAMETER: UninitializedParameter,
    _capi.NEGATIVE_DEGREES_OF_FREEDOM: NegativeDegreesOfFreedom,
    _capi.NO_FITTING_PARAMETERS: NoFittingParameters,
    _capi.NO_GLOBAL_PARAMETERS: NoGlobalParameters,
    _capi.INSUFFICIENT_INTEGRATION_WORKSPACE: InsufficientIntegrationWorkspace,
    _capi.UNKNOWN_MODEL: UnknownModel,
    _capi.CUDA_ERROR: CudaError,
    _capi.NCCL_ERROR: NcclError,
    _capi.INVALID_ARGUMENT: InvalidArgument,
    _capi.NOT_AVAILABLE: NotAvailable,
    _capi.TOO_MANY_PARAMETERS: TooManyParameters,
}


class Loss:
    """reference lm_solver.h:77-82"""

    linear = 0
    cauchy = 1
    huber = 2


class io:
    """IO flags, reference lm_solver.h:60-72"""

    all = 0b0000_0001
    delta1 = 0b0000_0010
    delta2 = 0b0000_0100
    timings = 0b0000_1000
    hide_local = 0b0001_0000
    hide_global = 0b0010_0000
    hide_all = 0b0100_0000
    final_only = 0b1000_0000


@dataclass
class SolverSettings:
    """``LMsolver::settings`` with the reference defaults (lm_solver.h:88-96)."""

    iteration_limit: int = 1000
    lambda_incs: int = 3
    lambda_down: float = 10.0
    lambda_up: float = 10.0
    damp_max: bool = True
    DTD_min: list = field(default_factory=lambda: [0.0])
    acceleration_threshold: float = -1.0
    verbosity: int = 0
    loss: int = Loss.linear
    n_threads: int = 1


@dataclass
class GadfOptions:
    """The arguments that exist only in the Fortran driver gadf_fit
    (fortran/gadfit/gadfit.F90:502-511).  enabled=True makes fit() follow
    gadf_fit (gadfit.F90:678-915) instead of LMsolver::fit; None = argument
    absent."""
    enabled: bool = False
    uphill: int = 0
    nielsen: bool = False
    umnigh: bool = False
    chi2_abs: float | None = None
    chi2_rel: float | None = None
    grad_chi2: float | None = None
    cos_phi: float | None = None
    rel_error: float | None = None
    rel_error_global: float | None = None


STOP_REASONS = ("iteration_limit", "lambda_incs", "chi2_abs", "chi2_rel", "grad_chi2", "cos_phi",
                "rel_error", "rel_error_global", "non_finite_chi2")


def _as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class LMsolver:
    GLOBAL = -1  # global_dataset_idx, lm_solver.h:99

    def __init__(self, model: str, _api: Api | None = None):
        self._api = _api if _api is not None else _capi.product_api()
        self._h = C.c_void_p()
        self._keep: list = []  # borrowed arrays must outlive fit()
        self.settings = SolverSettings()
        self.gadf = GadfOptions()
        self._check(self._api.create(model.encode(), C.byref(self._h)))

    # -- plumbing ---------------------------------------------------------
    def _check(self, status: int) -> None:
        if status == _capi.OK:
            return
        msg = ""
        if self._h:
            raw = self._api.last_error(self._h)
            msg = raw.decode() if raw else ""
        raise _STATUS_TO_EXCEPTION.get(status, GADfitException)(
            f"status {status}: {msg}")

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._api.destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _push_settings(self) -> None:
        s = Settings()
        st = self.settings
        s.iteration_limit = int(st.iteration_limit)
        s.lambda_incs = int(st.lambda_incs)
        s.lambda_down = float(st.lambda_down)
        s.lambda_up = float(st.lambda_up)
        s.damp_max = int(bool(st.damp_max))
        dtd = _as_f64(st.DTD_min)
        self._dtd_keep = dtd
        s.DTD_min = dtd.ctypes.data_as(_capi.c_double_p)
        s.n_DTD_min = int(dtd.size)
        s.acceleration_threshold = float(st.acceleration_threshold)
        s.verbosity = int(st.verbosity)
        s.loss = int(st.loss)
        s.n_threads = int(st.n_threads)
        self._check(self._api.set_settings(self._h, C.byref(s)))

    # -- reference interface ---------------------------------------------
    def addDataset(self, x_data, y_data=None, errors=None, n_datapoints=None):
        """lm_solver.h:158-167.  Arrays are borrowed (kept alive here).

        ``x_data=None`` with ``n_datapoints`` declares a data set that lives on
        another rank (multi-GPU only)."""
        if x_data is None:
            self._check(self._api.add_dataset(
                self._h, int(n_datapoints), None, None, None))
            return
        x = _as_f64(x_data)
        y = _as_f64(y_data)
        assert x.shape == y.shape and x.ndim == 1
        e = None
        if errors is not None:
            e = _as_f64(errors)
            assert e.shape == x.shape
        self._keep.append((x, y, e))
        n = x.size if n_datapoints is None else int(n_datapoints)
        self._check(self._api.add_dataset(
            self._h, n, x.ctypes.data, y.ctypes.data,
            e.ctypes.data if e is not None else None))

    def addDatasetPtr(self, n_datapoints: int, x_ptr: int, y_ptr: int,
                      errors_ptr: int | None = None):
        """Low level version taking raw host addresses (lm_solver.h:164-167);
        used with pinned torch buffers by bench.py."""
        self._check(self._api.add_dataset(
            self._h, int(n_datapoints), x_ptr, y_ptr, errors_ptr))

    def setPar(self, i_par: int, val: float, active: bool = False,
               i_dataset=GLOBAL, parameter_name: str = ""):
        """lm_solver.h:171-180; ``setPar(i, val, active, "name")`` sets a named
        global parameter like the second reference overload."""
        if isinstance(i_dataset, str):
            parameter_name, i_dataset = i_dataset, self.GLOBAL
        self._check(self._api.set_par(
            self._h, int(i_par), float(val), int(bool(active)), int(i_dataset),
            parameter_name.encode() if parameter_name else None))

    def _push_gadf_options(self):
        g = self.gadf
        if not self._api.has_gadf:
            if g.enabled:
                raise NotAvailable("this library has no gadf_fit driver")
            return
        absent = lambda v: -1.0 if v is None else float(v)  # noqa: E731
        opts = _capi.GadfOptions(int(g.enabled), int(g.uphill), int(g.nielsen), int(g.umnigh),
                                 absent(g.chi2_abs), absent(g.chi2_rel), absent(g.grad_chi2),
                                 absent(g.cos_phi), absent(g.rel_error), absent(g.rel_error_global))
        self._check(self._api.set_gadf_options(self._h, C.byref(opts)))

    def setLog(self, sink) -> None:
        """Text output of the reference's fit() -- iteration results, timing
        table, termination messages (lm_solver.cpp:485,498,587-680) -- one call
        of sink(line) per line, selected by settings.verbosity as in the
        reference.  None removes the sink (the default: nothing is printed)."""
        if not self._api.product:
            raise NotAvailable("setLog is a product entry point (the reference logs via spdlog)")
        if sink is None:
            self._log_cb = None
            self._check(self._api.set_log(self._h, _capi.LOG_FN(0), None))
            return
        self._log_cb = _capi.LOG_FN(lambda line, _user: sink(line.decode()))
        self._check(self._api.set_log(self._h, self._log_cb, None))

    def referenceLog(self) -> list[str]:
        """(compiled reference only) its own info-level output of the last fit"""
        need = C.c_size_t()
        self._check(self._api.get_log(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        self._check(self._api.get_log(self._h, buf, need.value, C.byref(need)))
        return buf.value.decode().split("\n")[:-1]

    def fit(self, lambda_: float = 10.0):
        self._push_settings()
        self._push_gadf_options()
        self._check(self._api.fit(self._h, float(lambda_)))

    def stopReason(self) -> str:
        """Which criterion ended the last fit (gadfit_b200_stop_reason)."""
        out = C.c_int()
        self._check(self._api.get_stop_reason(self._h, C.byref(out)))
        return STOP_REASONS[out.value]

    def getParValue(self, i_par: int, i_dataset: int = 0) -> float:
        out = C.c_double()
        self._check(self._api.get_par(self._h, i_par, i_dataset, C.byref(out)))
        return out.value

    def getValue(self, arg: float, i_dataset: int = 0) -> float:
        out = C.c_double()
        self._check(self._api.get_value(self._h, float(arg), i_dataset,
                                        C.byref(out)))
        return out.value

    def chi2(self) -> float:
        out = C.c_double()
        self._check(self._api.chi2(self._h, C.byref(out)))
        return out.value

    def degreesOfFreedom(self) -> int:
        out = C.c_int()
        self._check(self._api.dof(self._h, C.byref(out)))
        return out.value

    def _array(self, which: int) -> np.ndarray:
        need = C.c_size_t()
        self._check(self._api.get_array(self._h, which, None, 0, C.byref(need)))
        buf = np.empty(need.value, dtype=np.float64)
        if need.value:
            self._check(self._api.get_array(
                self._h, which, buf.ctypes.data, buf.size, C.byref(need)))
        return buf

    def getJacobian(self):
        return self._array(_capi.JACOBIAN)

    def getJTJ(self):
        return self._array(_capi.JTJ)

    def getDTD(self):
        return self._array(_capi.DTD)

    def getLeftSide(self):
        return self._array(_capi.LEFT_SIDE)

    def getRightSide(self):
        return self._array(_capi.RIGHT_SIDE)

    def getResiduals(self):
        return self._array(_capi.RESIDUALS)

    def getInvJTJ(self):
        return self._array(_capi.INV_JTJ)

    def getInvJTJDiagonal(self):
        """diag((J^T J)^-1), one entry per Jacobian column: the parameter
        variances up to chi2/dof.  Computed from the block-arrow factors, so
        available for any number of data sets (the dense getInvJTJ() is
        limited to n^2 <= 2^27 entries)."""
        return self._array(_capi.INV_JTJ_DIAGONAL)

    def getDelta1(self):
        return self._array(_capi.DELTA1)

    def getDelta2(self):
        return self._array(_capi.DELTA2)

    def getOmega(self):
        return self._array(_capi.OMEGA)

    # -- additions over the reference --------------------------------------
    def trace(self) -> list[dict]:
        """Accept/reject history of the last fit: what the reference prints as
        "Current lambda: {}, new chi2 = {}, old chi2: {}" (lm_solver.cpp:457)."""
        n = C.c_size_t()
        self._check(self._api.get_trace(self._h, None, 0, C.byref(n)))
        buf = (Trial * max(n.value, 1))()
        self._check(self._api.get_trace(self._h, buf, n.value, C.byref(n)))
        return [
            dict(lambda_=t.lambda_, new_chi2=t.new_chi2, old_chi2=t.old_chi2,
                 iteration=t.iteration, decision=t.decision)
            for t in buf[: n.value]
        ]

    def iterations(self) -> int:
        out = C.c_int()
        self._check(self._api.n_iterations(self._h, C.byref(out)))
        return out.value

    def timings(self) -> dict:
        t = Timings()
        self._check(self._api.get_timings(self._h, C.byref(t)))
        names = ["jacobian", "chi2", "linalg", "omega", "main"]
        return {k: (t.seconds[i], t.calls[i]) for i, k in enumerate(names)}

    # -- product-only ------------------------------------------------------
    def launches(self) -> int:
        out = C.c_longlong()
        self._check(self._api.n_launches(self._h, C.byref(out)))
        return out.value

    def jacobianKernelTime(self) -> tuple[float, int]:
        sec = C.c_double()
        n = C.c_int()
        self._check(self._api.jacobian_kernel_time(self._h, C.byref(sec),
                                                   C.byref(n)))
        return sec.value, n.value

    def commStats(self) -> dict:
        """Multi-GPU: peer-memory all-reduce active? and reduction counts."""
        active, peer, nccl = C.c_int(), C.c_longlong(), C.c_longlong()
        self._check(self._api.comm_stats(self._h, C.byref(active), C.byref(peer),
                                         C.byref(nccl)))
        return dict(peer_reduce_active=bool(active.value), peer_reductions=peer.value,
                    nccl_reductions=nccl.value)

    def setDevice(self, device: int):
        self._check(self._api.set_device(self._h, int(device)))

    def upload(self):
        self._check(self._api.upload(self._h))

    def invalidateData(self):
        """The data arrays were changed in place: copy them to HBM again at the
        next fit() / chi2() / upload().  (The reference re-reads the caller's
        arrays at every fit(), lm_solver.h:105-113.)"""
        self._check(self._api.invalidate_data(self._h))

    def setPhaseTimers(self, on: bool):
        """The per-phase CUDA-event timers (timings(), jacobianKernelTime()) are
        on by default; off saves twelve event records per trial step, which
        shows when a trial step takes ~100 us (gadfit_b200_set_phase_timers)."""
        self._check(self._api.set_phase_timers(self._h, 1 if on else 0))

    def setComm(self, rank: int, world_size: int, unique_id: bytes):
        assert len(unique_id) == 128
        self._id_keep = C.create_string_buffer(unique_id, 128)
        self._check(self._api.set_comm(self._h, rank, world_size,
                                       self._id_keep))


def comm_unique_id() -> bytes:
    """128-byte NCCL unique id (rank 0 creates it, the caller broadcasts it)."""
    api = _capi.product_api()
    buf = C.create_string_buffer(128)
    status = api.comm_unique_id(buf)
    if status != _capi.OK:
        raise NcclError(f"status {status}")
    return buf.raw


def partition(n_points: list[int], rank: int, world_size: int) -> tuple[int, int, bool]:
    """Which data sets ``rank`` owns: (first, end, split) -- see
    ``gadfit_b200_partition``."""
    api = _capi.product_api()
    arr = (C.c_int * len(n_points))(*n_points)
    first, end, split = C.c_int(), C.c_int(), C.c_int()
    status = api.partition(len(n_points), arr, rank, world_size,
                           C.byref(first), C.byref(end), C.byref(split))
    if status != _capi.OK:
        raise InvalidArgument(f"status {status}")
    return first.value, end.value, bool(split.value)


def measure_fp64_peak() -> float:
    api = _capi.product_api()
    out = C.c_double()
    status = api.measure_fp64_peak(C.byref(out))
    if status != _capi.OK:
        raise CudaError(f"status {status}")
    return out.value


def load_models(path: str) -> None:
    """Load a model plug-in (include/gadfit_b200/model_plugin.h): user-defined
    fitting functions compiled by `python -m gadfit_b200.build --plugin`; they
    become available to LMsolver(name).  The counterpart of handing the
    reference an arbitrary lambda (lm_solver.h:146-148)."""
    status = _capi.product_api().load_models(os.fspath(path).encode())
    if status != _capi.OK:
        raise InvalidArgument(
            f"{path}: cannot be loaded, is not a model plug-in built against these "
            f"headers, or registers a name that is already taken")


def load_test_models() -> None:
    """The plug-in with the functions of the reference's nested and direct
    double-integral tests (include/gadfit_b200/test_models.h)."""
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(here, "libgadfit_b200_testmodels.so")
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    load_models(path)


def models() -> dict[str, int]:
    api = _capi.product_api()
    out = {}
    for i in range(api.n_models()):
        name = C.c_char_p()
        n = C.c_int()
        api.model_info(i, C.byref(name), C.byref(n))
        out[name.value.decode()] = n.value
    return out
