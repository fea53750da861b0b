"""User-defined fitting functions through model plug-ins
(include/gadfit_b200/model_plugin.h, gadfit_b200_load_models): the counterpart
of handing the reference's LMsolver an arbitrary lambda (lm_solver.h:146-148,
c++/tests/example.cpp:11-19)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import backends
import scenarios
import gadfit_b200 as g
from gadfit_b200 import build as gbuild

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "user_models_plugin.cu")
OUT = os.path.join(HERE, "_build", "libuser_models.so")


@pytest.fixture(scope="module")
def user_plugin():
    """Built with the documented command line; the artefact travels to the GPU
    box with the snapshot, so it is only rebuilt when stale."""
    deps = [SRC, os.path.join(gbuild.ROOT, "include", "gadfit_b200", "model_plugin.h"),
            os.path.join(gbuild.CSRC, "sweep_kernels.cuh"), os.path.join(gbuild.CSRC, "model_ops.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run([sys.executable, "-m", "gadfit_b200.build", "--plugin", SRC, "-o", OUT],
                       check=True, cwd=gbuild.ROOT)
    g.load_models(OUT)
    return OUT


def test_plugin_registers_its_functions_without_a_gpu(user_plugin):
    registered = g.models()
    assert registered["user_decay"] == 3
    assert registered["user_stretched_integral"] == 4
    assert all(registered[f"user_stretched_integral_k{p}"] == 4 for p in (21, 31, 41, 51, 61))
    g.load_models(user_plugin)  # loading twice is a no-op
    assert list(g.models()).count("user_decay") == 1
    s = g.LMsolver("user_decay")  # host-side bookkeeping works without a device
    s.addDataset(np.arange(4.0), np.arange(4.0))
    s.setPar(0, 1.0, True)
    with pytest.raises(g.UnknownModel):
        g.LMsolver("user_decay_typo")


def test_loading_something_that_is_not_a_plugin_fails_cleanly(tmp_path):
    with pytest.raises(g.InvalidArgument):
        g.load_models(str(tmp_path / "does_not_exist.so"))
    oracle_lib = backends.oracle_api().path  # a shared library, but no plug-in
    with pytest.raises(g.InvalidArgument):
        g.load_models(oracle_lib)


def test_test_model_plugin_provides_the_integration_test_functions():
    g.load_test_models()
    registered = g.models()
    assert sum(k.startswith("nested_integral_") for k in registered) == 12
    assert sum(k.startswith("direct_integral_") for k in registered) == 15
    assert registered["direct_integral_y1y2x1x2"] == 7 and registered["nested_integral_x1x2"] == 6


@pytest.mark.gpu
def test_user_function_reproduces_the_builtin_known_answers(user_plugin, c1):
    """The same functor body, compiled in a plug-in under a user's name, gives
    bit-identical results to the built-in one and hence the reference's known
    answers (c++/tests/lm_solver.cpp:21-47)."""
    sc = scenarios.C1_SCENARIOS[0]
    builtin, _ = scenarios.run(sc, lambda _m: g.LMsolver("exponential"), c1)
    user, _ = scenarios.run(sc, lambda _m: g.LMsolver("user_decay"), c1)
    assert builtin["chi2"] == user["chi2"] and builtin["pars"] == user["pars"]
    assert [t["new_chi2"] for t in builtin["trace"]] == [t["new_chi2"] for t in user["trace"]]
    np.testing.assert_allclose(user["chi2"], sc["kat"][0], rtol=1e-13)


@pytest.mark.gpu
def test_user_function_with_adaptive_integral_recovers_its_parameters(user_plugin):
    """A function the library has never seen (adaptive integral with the data
    point as upper bound, pow/exp inside): fit noise-free synthetic data."""
    from scipy.integrate import quad
    truth = (2.0, 1.5, 0.8, 0.3)
    x = np.linspace(0.2, 6.0, 120)
    y = np.array([truth[0] * quad(lambda t: np.exp(-(t / truth[1]) ** truth[2]), 0.0, xi,
                                  epsabs=1e-13, epsrel=1e-13)[0] + truth[3] for xi in x])
    s = g.LMsolver("user_stretched_integral")
    s.addDataset(x, y)
    for k, v in enumerate((1.5, 1.2, 1.0, 0.1)):
        s.setPar(k, v, True)
    s.settings.iteration_limit = 50
    s.fit(1.0)
    got = [s.getParValue(k) for k in range(4)]
    np.testing.assert_allclose(got, truth, rtol=1e-6)
    assert s.chi2() < 1e-14
    np.testing.assert_allclose(s.getValue(3.0), truth[0] * quad(
        lambda t: np.exp(-(t / truth[1]) ** truth[2]), 0.0, 3.0)[0] + truth[3], rtol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("pts", (21, 31, 41, 51, 61))
def test_other_gauss_kronrod_rules_on_the_device(user_plugin, pts):
    """The Fortran version's integration rules (set_integration_rule) as a
    template argument of c.integrate: the same fit as with the default 15-point
    rule -- every rule integrates adaptively to 1e-10, so the fitted parameters
    agree far below the data's information content."""
    truth = (2.0, 1.5, 0.8, 0.3)
    rng = np.random.default_rng(7)
    x = np.linspace(0.2, 6.0, 120)
    from scipy.integrate import quad
    y = np.array([truth[0] * quad(lambda t: np.exp(-(t / truth[1]) ** truth[2]), 0.0, xi,
                                  epsabs=1e-13, epsrel=1e-13)[0] + truth[3] for xi in x])
    y = y + 1e-3 * rng.standard_normal(x.size)

    def fit(model):
        s = g.LMsolver(model)
        s.addDataset(x, y)
        for k, v in enumerate((1.5, 1.2, 1.0, 0.1)):
            s.setPar(k, v, True)
        s.settings.iteration_limit = 12
        s.settings.acceleration_threshold = 0.9  # omega sweep: second order through the rule
        s.fit(1.0)
        return np.array([s.getParValue(k) for k in range(4)] + [s.chi2()]), s.trace()

    want, want_trace = fit("user_stretched_integral")
    got, got_trace = fit(f"user_stretched_integral_k{pts}")
    np.testing.assert_allclose(got, want, rtol=1e-8)
    # identical decisions until chi2 has converged (afterwards accept / reject
    # is decided by the last bits, cf. c++/tests/lm_solver.cpp:396-398)
    assert [t["decision"] for t in got_trace[:8]] == [t["decision"] for t in want_trace[:8]]


# ---- fitting functions defined from Python (gadfit_b200.define_model) -------

def test_define_model_generates_compiles_and_registers(tmp_path):
    src = g.model_source("py_decay", 3, "P(0) * exp(-x / P(1)) + P(2)")
    assert "GADFIT_B200_REGISTER_MODEL(UserModel_py_decay)" in src and "return P(0)" in src
    so = g.define_model("py_decay", 3, "P(0) * exp(-x / P(1)) + P(2)", cache_dir=str(tmp_path))
    assert os.path.exists(so) and g.models()["py_decay"] == 3
    with pytest.raises(g.InvalidArgument):
        g.define_model("py_decay", 3, "P(0)", cache_dir=str(tmp_path))  # name taken
    with pytest.raises(g.InvalidArgument):
        g.model_source("not an identifier", 1, "P(0)")


@pytest.mark.gpu
def test_python_defined_function_reproduces_the_builtin_known_answers(c1, tmp_path):
    if "py_decay_gpu" not in g.models():
        g.define_model("py_decay_gpu", 3, "P(0) * exp(-x / P(1)) + P(2)", cache_dir=str(tmp_path))
    sc = scenarios.C1_SCENARIOS[0]
    builtin, _ = scenarios.run(sc, lambda _m: g.LMsolver("exponential"), c1)
    user, _ = scenarios.run(sc, lambda _m: g.LMsolver("py_decay_gpu"), c1)
    assert builtin["chi2"] == user["chi2"] and builtin["pars"] == user["pars"]
