"""Text output of fit(): printIterationResults / printTimings and the
termination messages of the reference (lm_solver.cpp:485,498,587-680), emitted
by the library through gadfit_b200_set_log and selected by settings.verbosity
(io::flag bits, lm_solver.h:60-72).

The expected text is the unmodified reference's own: oracle/spdlog_stub formats
its info-level messages with std::vformat (the semantics of the fmt library
spdlog uses), oracle/ref_harness.cpp hands them out.  Lines must match token
for token; numbers to 1e-9 (they are printed with 15-17 digits)."""
import re

import numpy as np
import pytest

import backends
import gadfit_b200 as g
import scenarios

IO = dict(all=1, delta1=2, delta2=4, timings=8, hide_local=16, hide_global=32, hide_all=64,
          final_only=128)
NUMBER = re.compile(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?|nan|inf")

pytestmark = pytest.mark.skipif(not backends.has_ref(), reason="oracle/_ref not built")


def fit_c1(make, c1, verbosity, named, settings=None, sink=None, fixed=False):
    sc = dict(scenarios.C1_SCENARIOS[0])
    sc["settings"] = dict(sc["settings"]) | dict(acceleration_threshold=0.9) | (settings or {})
    s, lam, _, _ = scenarios.build(sc, make, c1)
    if named:  # lm_solver.cpp:99-104: names are per parameter index
        d = c1["fix_d"]
        s.setPar(0, d[scenarios.I0_0], True, 0, "I0")
        s.setPar(1, d[scenarios.TAU], True, -1, "tau")
    if fixed:  # " (fixed)" instead of the step, lm_solver.cpp:621-623
        s.setPar(2, c1["fix_d"][scenarios.BGR_1], False, 1)
    s.settings.verbosity = verbosity | IO["timings"]  # the harness always asks for the table
    if sink is not None:
        s.setLog(sink)
    s.fit(lam)
    return s


def compare(got, want):
    assert len(got) == len(want), "\n".join(got) + "\n--- reference ---\n" + "\n".join(want)
    in_table = False
    for a, b in zip(got, want):
        if a.startswith("Timings"):
            in_table = True
        skeleton_a, skeleton_b = NUMBER.sub("#", a), NUMBER.sub("#", b)
        na, nb = [float(t) for t in NUMBER.findall(a)], [float(t) for t in NUMBER.findall(b)]
        if in_table and nb:
            # seconds differ by construction; the label and, for the sweeps,
            # the number of calls must agree
            label = a[:14]
            assert label == b[:14]
            if label.strip() in ("Jacobian", "Chi2", "Omega", "Main loop"):
                assert na[-1] == nb[-1], (a, b)
            continue
        # widths of right-aligned numbers may differ with the digits printed
        assert re.sub(r"\s+", " ", skeleton_a) == re.sub(r"\s+", " ", skeleton_b), (a, b)
        np.testing.assert_allclose(na, nb, rtol=1e-9, atol=1e-12, err_msg=f"{a!r} vs {b!r}")


def test_reference_log_capture_has_the_expected_shape(c1):
    ref = fit_c1(backends.ref_solver(), c1, 0, named=True)
    log = ref.referenceLog()
    assert log[0] == "Iteration: 1" and log[1].startswith("Lambda: ") and log[2].startswith("Chi2/DOF: ")
    assert "  Global parameters" in log and "  Data set: 1" in log
    assert any(line.startswith("            tau: ") for line in log)
    assert any(line.startswith("    Parameter 2: ") for line in log)
    assert "Iteration limit reached" in log
    assert "Timings          Wall (s)    CPU (s)  CPU rel  Calls" in log


CASES = [
    ("every_iteration", 0, True, None),
    ("unnamed", 0, False, None),
    ("with_deltas", IO["delta1"] | IO["delta2"], True, None),
    ("all", IO["all"], True, None),
    ("final_only", IO["final_only"], True, None),
    ("hide_local", IO["hide_local"], True, None),
    ("hide_global", IO["hide_global"] | IO["delta1"], True, None),
    ("hide_all", IO["hide_all"], True, None),
    ("no_iterations", 0, True, dict(iteration_limit=0)),
    ("gives_up", 0, True, dict(iteration_limit=60, lambda_incs=1)),
    ("one_fixed_parameter", IO["delta1"], True, "fixed"),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_device_text_output_matches_the_reference(case, c1):
    _, verbosity, named, settings = case
    fixed = settings == "fixed"
    settings = None if fixed else settings
    lines = []
    solvers = []
    for make, sink in ((backends.product_solver(), lines.append), (backends.ref_solver(), None)):
        solvers.append(fit_c1(make, c1, verbosity, named, settings, sink, fixed))
    compare(lines, solvers[1].referenceLog())


@pytest.mark.gpu
def test_no_sink_no_output_and_sink_can_be_removed(c1, capfd):
    s = fit_c1(backends.product_solver(), c1, 0, named=True)
    lines = []
    s.setLog(lines.append)
    s.fit(1.0)
    assert lines and lines[0] == "Iteration: 1"
    n = len(lines)
    s.setLog(None)
    s.fit(1.0)
    assert len(lines) == n
    out = capfd.readouterr()
    assert out.out == ""
