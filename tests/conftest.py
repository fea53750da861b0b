import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
    config.addinivalue_line("markers", "slow: full-size runs of the CPU reference beside the GPU (about a minute)")


@pytest.fixture(scope="session")
def c1():
    """Data of the reference's LM-solver tests (c++/tests/fixtures.h:9-68)."""
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "c1_fixture.json")))
    return {k: (np.array(v) if isinstance(v, list) else v) for k, v in fx.items()}


@pytest.fixture(scope="session")
def golden():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "golden_ref.json")))


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Which tolerance every trace / parameter comparison of this session met
    (tests/scenarios.py: compare_to_golden)."""
    import scenarios
    log = scenarios.PARITY_LOG
    if not log:
        return
    tr = terminalreporter
    tr.section("parity of accept/reject traces and parameters")
    relaxed = [r for r in log if r["needed_relaxed"]]
    to_end = sum(r["trace_identical_to_end"] for r in log)
    tr.write_line(f"{len(log)} comparisons: {to_end} with a trace identical to the last trial, "
                  f"{len(log) - len(relaxed)} with parameters within {scenarios.STRICT_PAR_TOL:g}, "
                  f"{len(relaxed)} needed the post-convergence tolerance {scenarios.RELAXED_PAR_TOL:g}")
    worst = max(log, key=lambda r: r["worst_parameter_error"])
    tr.write_line(f"worst parameter error {worst['worst_parameter_error']:.2e} ({worst['name']} {worst['label']}), "
                  f"worst chi2 error {max(r['chi2_error'] for r in log):.2e}")
    for r in log:
        if r["needed_relaxed"] or not r["trace_identical_to_end"]:
            tr.write_line(
                f"  {r['name']:<34} {r['label']:<8} trials {r['trials_identical']}/{r['trials']} identical "
                f"(compared {r['trials_compared']}), parameters {r['worst_parameter_error']:.1e} "
                f"<= {r['parameter_tolerance']:g}" + ("  RELAXED" if r["needed_relaxed"] else ""))
    out = os.environ.get("GADFIT_PARITY_LOG")
    if out:
        with open(out, "w") as f:
            json.dump(log, f, indent=1)
