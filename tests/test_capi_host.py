"""CPU: the C-ABI library loads, exports every symbol include/gadfit_b200.h
declares, and its host-side bookkeeping (the reference's call-order rules and
index map) works without a GPU.  No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gadfit_b200 as g
from gadfit_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "gadfit_b200.h")).read()
    declared = set(re.findall(r"GADFIT_B200_API\s+(?:const char\*|int)\s+(gadfit_b200_\w+)\(", header))
    assert len(declared) >= 28
    lib = C.CDLL(_capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared but not exported"
    bound = {"gadfit_b200_" + n for n in _capi.CORE_SYMBOLS + _capi.PRODUCT_ONLY_SYMBOLS}
    assert bound == declared, bound ^ declared


def test_model_registry():
    registered = g.models()
    for name, n_pars in {"exponential": 3, "gauss_hermite3": 20, "gauss_linear_bg": 6,
                         "conv_exp_gauss": 5, "lorentz_double_int": 7}.items():
        assert registered[name] == n_pars
    singles = {k: v for k, v in registered.items() if k.startswith("single_integral_")}
    assert len(singles) == 8 and set(singles.values()) == {2}
    with pytest.raises(g.UnknownModel):
        g.LMsolver("no_such_function")


def test_call_order_rules_without_a_gpu(c1):
    """Reference lm_solver.cpp:35-37,65-68,119-190: these are decided on the
    host before any device work, so they hold on a machine without a GPU."""
    from test_oracle import _run_exception_sections
    _run_exception_sections(lambda m: g.LMsolver(m), c1, g)


def test_default_settings_are_the_reference_defaults():
    """lm_solver.h:88-96"""
    s = _capi.Settings()
    assert _capi.product_api().default_settings(C.byref(s)) == 0
    assert (s.iteration_limit, s.lambda_incs, s.lambda_down, s.lambda_up, s.damp_max,
            s.acceleration_threshold, s.loss, s.n_threads) == (1000, 3, 10.0, 10.0, 1, -1.0, 0, 1)
    st = g.SolverSettings()
    assert (st.iteration_limit, st.lambda_incs, st.damp_max, st.DTD_min) == (1000, 3, True, [0.0])


def test_partition_whole_data_sets_balanced_by_points():
    """Multi-GPU sharding: contiguous blocks of whole data
    sets, balanced by points; fewer sets than ranks -> split points."""
    n = [1000] * 64
    blocks = [g.partition(n, r, 8) for r in range(8)]
    assert blocks == [(8 * r, 8 * r + 8, False) for r in range(8)]
    # uneven sizes: every data set has exactly one owner, blocks are contiguous
    rng = np.random.default_rng(0)
    n = rng.integers(1, 5000, size=1000).tolist()
    blocks = [g.partition(n, r, 8) for r in range(8)]
    assert blocks[0][0] == 0 and blocks[-1][1] == 1000
    for a, b in zip(blocks, blocks[1:]):
        assert a[1] == b[0]
    loads = [sum(n[a:b]) for a, b, _ in blocks]
    assert max(loads) < 1.05 * sum(n) / 8
    assert g.partition([10, 20, 30], 1, 4) == (0, 3, True)
    assert g.partition([10, 20, 30], 0, 1) == (0, 3, False)


def test_parameters_are_readable_before_fit(c1):
    s = g.LMsolver("exponential")
    s.addDataset(c1["x_data_1"], c1["y_data_1"])
    s.setPar(0, 1.5, True, 0)
    s.setPar(1, 2.5, True)
    s.setPar(2, 3.5, False, 0)
    assert [s.getParValue(k) for k in range(3)] == [1.5, 2.5, 3.5]


def test_host_only_switches_work_without_a_gpu(c1):
    """Round 2 entry points that touch no device state: the phase-timer switch
    and invalidate_data, on a machine without a GPU; null handles are refused."""
    s = g.LMsolver("exponential")
    s.addDataset(c1["x_data_1"], c1["y_data_1"])
    s.setPhaseTimers(False)
    s.setPhaseTimers(True)
    s.invalidateData()
    api = _capi.product_api()
    assert api.set_phase_timers(None, 1) != _capi.OK
    assert api.invalidate_data(None) != _capi.OK
