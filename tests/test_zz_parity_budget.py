"""Runs last (file name): how many of this session's trace / parameter
comparisons (scenarios.compare_to_golden) NEEDED the relaxed post-convergence
parameter tolerance of 1e-7 instead of the north star's 1e-10.  The relaxation
is argued by the reference's own test comment (c++/tests/lm_solver.cpp:396-401:
after convergence the decisions are rounding noise); this test keeps it an
exception: at most 1 comparison in 10 may need it, and every comparison whose
trace is identical to the end must meet 1e-10 (asserted inside
compare_to_golden)."""
import pytest

import scenarios


def _check():
    log = scenarios.PARITY_LOG
    if not log:
        pytest.skip("no trace comparisons ran in this session")
    relaxed = [r for r in log if r["needed_relaxed"]]
    assert len(relaxed) <= max(1, len(log) // 10), [
        (r["name"], r["label"], r["worst_parameter_error"]) for r in relaxed]
    for r in log:
        if r["trace_identical_to_end"]:
            assert not r["needed_relaxed"], r


def test_relaxed_parameter_tolerance_is_the_exception_cpu():
    _check()


@pytest.mark.gpu
def test_relaxed_parameter_tolerance_is_the_exception_gpu():
    _check()
