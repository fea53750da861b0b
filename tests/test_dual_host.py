"""CPU: the product's dual numbers (include/gadfit_b200/dual.h, the same
header the kernels compile) against the reference's AD known-answer values:
forward mode c++/tests/ad_forward_mode.cpp:16-256 (val, d, dd @1e-13) and the
reverse-mode gradients c++/tests/ad_reverse_mode.cpp (@1e-13 here)."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# name -> (val, d, dd) asserted by ad_forward_mode.cpp
FORWARD_KAT = {
    "arith_none": (67.7091669953527, 0.0, 0.0),
    "arith_a": (67.7091669953527, 0.3757755919598275, 0.4048404006248003),
    "arith_b": (67.7091669953527, 11.28454180719661, 10.33354861677357),
    "arith_c": (67.7091669953527, -4.368251216348313, -4.368251216348313),
    "arith_ab": (67.7091669953527, 11.66031739915644, 10.68025940006843),
    "arith_ac": (67.7091669953527, -3.992475624388486, -3.936039790229344),
    "arith_bc": (67.7091669953527, 6.916290590848295, 3.937926374931089),
    "arith_abc": (67.7091669953527, 7.292066182808123, 4.312008183720113),
    "exp_none": (402.2477537977381, 0.0, 0.0),
    "exp_a": (402.2477537977381, -4.467047498107922, -1.74515271385656),
    "exp_b": (402.2477537977381, 387.314505060867, 672.8348977983287),
    "exp_ab": (402.2477537977381, 382.8474575627591, 670.5530084135362),
    "trig_none": (-0.5540770421819348, 0.0, 0.0),
    "trig_a": (-0.5540770421819348, -1.549501027521998, -19.82580971358727),
    "trig_b": (-0.5540770421819348, 0.4556944655440529, 1.648782695748266),
    "trig_ab": (-0.5540770421819348, -1.093806561977945, -16.07874055136707),
    "erf_none": (0.5512884666654083, 0.0, 0.0),
    "erf_a": (0.5512884666654083, 0.8469022413858851, -0.06043365341217193),
}
# gradients: the first derivatives above with a single active variable are
# exactly what ad_reverse_mode.cpp:15-17 asserts as adjoints (a_ref, b_ref, c_ref)
GRADIENT_KAT = {
    "grad_arith": (67.7091669953527, 0.3757755919598275, 11.28454180719661, -4.368251216348313),
    "grad_exp": (402.2477537977381, -4.467047498107922, 387.314505060867, 0.0),
    "grad_trig": (-0.5540770421819348, -1.549501027521998, 0.4556944655440529, 0.0),
    "grad_erf": (0.5512884666654083, 0.8469022413858851, 0.0, 0.0),
    "grad_arith_b_only": (67.7091669953527, 0.0, 11.28454180719661, 0.0),
}


# every elemental operation on its own, evaluated by the reference's own AdVar
# in forward mode (tests/golden/make_dual_golden.py)
SINGLE_OPS = json.load(open(os.path.join(ROOT, "tests", "golden", "dual_ops_ref.json")))["values"]


@pytest.fixture(scope="module")
def host_values(tmp_path_factory):
    exe = tmp_path_factory.mktemp("dual") / "host_dual_check"
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.run(["g++", "-std=c++20", "-O2", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "host_dual_check.cpp"), "-o", str(exe)],
                   check=True, env=env)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    return {ln.split()[0]: [float(v) for v in ln.split()[1:]] for ln in out.splitlines()}


@pytest.mark.parametrize("name", sorted(FORWARD_KAT))
def test_forward_mode_known_answers(name, host_values):
    np.testing.assert_allclose(host_values[name][:3], FORWARD_KAT[name], rtol=1e-13, atol=0.0)


@pytest.mark.parametrize("name", sorted(GRADIENT_KAT))
def test_static_sparsity_gradients(name, host_values):
    np.testing.assert_allclose(host_values[name], GRADIENT_KAT[name], rtol=1e-13, atol=0.0)


@pytest.mark.parametrize("name", sorted(SINGLE_OPS))
def test_single_operations_against_the_reference_advar(name, host_values):
    """All 25 elemental operations (12 trigonometric / hyperbolic, exp, log,
    sqrt, abs, erf, 4 pow variants, + - * /) one at a time: Dir2 (val, d, dd)
    and SDual gradients against gadfit::AdVar itself."""
    np.testing.assert_allclose(host_values[name][:3], SINGLE_OPS[name], rtol=1e-13, atol=1e-300)
