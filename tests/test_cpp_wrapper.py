"""The header-only C++ mirror of gadfit::LMsolver
(include/gadfit_b200/lm_solver.h) compiles with g++ against the C-ABI
library; its call-order rules work on the host; on a GPU the same program
reproduces the oracle's fit."""
import os
import subprocess

import numpy as np
import pytest

import backends

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("cpp") / "cpp_example"
    env = dict(os.environ)
    env.pop("CXX", None)
    lib_dir = os.path.join(ROOT, "gadfit_b200")
    subprocess.run(["g++", "-std=c++20", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp_example.cpp"), "-o", str(out),
                    "-L", lib_dir, "-lgadfit_b200", f"-Wl,-rpath,{lib_dir}"],
                   check=True, env=env)
    return str(out)


def test_cpp_wrapper_compiles_and_enforces_call_order(exe):
    plugin = os.path.join(ROOT, "gadfit_b200", "libgadfit_b200_testmodels.so")
    r = subprocess.run([exe, "host", plugin], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "caught 3" in r.stdout
    # gadfit::loadModels: user-defined functions from a model plug-in
    assert "plugin loaded" in r.stdout and "bad plugin rejected" in r.stdout


@pytest.mark.gpu
def test_cpp_wrapper_fit_matches_oracle(exe):
    r = subprocess.run([exe, "fit"], capture_output=True, text=True, check=True)
    out = r.stdout.splitlines()
    report = out[out.index("report_begin") + 1:out.index("report_end")]
    # gadfit::LMsolver prints like the reference does by default (stdout)
    assert report[0] == "Iteration limit reached" and report[1] == "Iteration: 1"
    assert "  Global parameters" in report and "  Data set: 1" in report
    got = {ln.split()[0]: ln.split()[1:] for ln in out[:out.index("report_begin")] if ln.strip()}
    x = np.arange(1.0, 9.0)
    y1 = np.array([56.0, 51.0, 54.0, 46.0, 50.0, 48.0, 31.0, 45.0])
    y2 = np.array([153.0, 157.0, 137.0, 111.0, 126.0, 108.0, 113.0, 128.0])
    s = backends.oracle_solver(1)("exponential")
    s.addDataset(x, y1)
    s.addDataset(x, y2)
    s.setPar(0, 50.0, True, 0)
    s.setPar(2, 1.0, True, 0)
    s.setPar(0, 150.0, True, 1)
    s.setPar(2, 5.5, False, 1)
    s.setPar(1, 10.0, True)
    s.settings.iteration_limit = 6
    s.settings.acceleration_threshold = 0.9
    s.fit(1.0)
    np.testing.assert_allclose(float(got["chi2"][0]), s.chi2(), rtol=1e-10)
    np.testing.assert_allclose(float(got["tau"][0]), s.getParValue(1), rtol=1e-10)
    np.testing.assert_allclose(float(got["I0_0"][0]), s.getParValue(0, 0), rtol=1e-10)
    np.testing.assert_allclose(float(got["I0_1"][0]), s.getParValue(0, 1), rtol=1e-10)
    assert int(got["iterations"][0]) == s.iterations()
    assert int(got["jtj_size"][0]) == 16
