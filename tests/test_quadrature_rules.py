"""CPU: the Gauss-Kronrod rules of include/gadfit_b200/quadrature.h -- the
15-point rule of the C++ reference and the 21 ... 61-point rules the Fortran
version offers (set_integration_rule, fortran/gadfit/numerical_integration.F90:
139-171).  The tables are generated from the rules' definition
(tools/kronrod_tables.py) and pinned on the reference's own tables
(fortran/gadfit/gauss_kronrod_parameters.F90) where the reference is mounted;
the header is compiled for the host here (it is the header the kernels compile)
and every rule is checked on its defining properties and on an integral with
AD-valued parameter and bound."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RULES = (15, 21, 31, 41, 51, 61)
REFERENCE_TABLES = "/root/reference/fortran/gadfit/gauss_kronrod_parameters.F90"


@pytest.fixture(scope="module")
def host_values(tmp_path_factory):
    exe = tmp_path_factory.mktemp("quadrature") / "host_quadrature_check"
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.run(["g++", "-std=c++20", "-O2", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "host_quadrature_check.cpp"), "-o", str(exe)],
                   check=True, env=env)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    return {ln.split()[0]: [float(v) for v in ln.split()[1:]] for ln in out.splitlines()}


@pytest.mark.parametrize("pts", RULES)
def test_rule_weights_and_polynomial_exactness(pts, host_values):
    """Kronrod with 2n+1 points is exact up to degree 3n+1, its Gauss rule up to 2n-1."""
    np.testing.assert_allclose(host_values[f"weights_{pts}"], [2.0, 2.0], rtol=5e-16)
    assert max(host_values[f"exactness_{pts}"]) < 2e-15


def _analytic(a, b):
    """F = int_0^b e^{-a t} cos(3 t) dt and its derivatives, by complex arithmetic:
    F = Re G, G = (1 - e^{-z b}) / z, z = a - 3i."""
    z = complex(a, -3.0)
    e = np.exp(-z * b)
    G = (1 - e) / z
    Ga = (b * e * z - (1 - e)) / z ** 2          # dG/da (dz/da = 1)
    Gb = e                                        # dG/db = integrand at b
    Gaa = (-b * b * e * z ** 2 - 2 * (b * e * z - (1 - e))) / z ** 3
    Gab = -b * e
    Gbb = -z * e
    return G.real, Ga.real, Gb.real, Gaa.real, Gab.real, Gbb.real


@pytest.mark.parametrize("pts", RULES)
def test_integral_with_ad_valued_parameter_and_bound(pts, host_values):
    """Fixed panel, adaptive integration (first derivatives, Leibniz term of
    the AD-valued upper bound) and the second directional derivative."""
    F, Fa, Fb, Faa, Fab, Fbb = _analytic(0.7, 2.3)
    fixed = host_values[f"fixed_{pts}"]
    np.testing.assert_allclose(fixed, [F, Fa], rtol=5e-10 if pts == 15 else 1e-13)
    np.testing.assert_allclose(host_values[f"adaptive_{pts}"], [F, Fa, Fb], rtol=2e-12)
    da, db = 1.0, 0.5
    d1 = Fa * da + Fb * db
    d2 = Faa * da * da + 2 * Fab * da * db + Fbb * db * db
    np.testing.assert_allclose(host_values[f"dir2_{pts}"], [F, d1, d2], rtol=2e-12)


def test_committed_tables_are_the_generators_output():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kronrod_tables.py")],
                         check=True, capture_output=True, text=True).stdout
    committed = open(os.path.join(ROOT, "include", "gadfit_b200", "quadrature_rules.h")).read()
    assert out == committed


@pytest.mark.skipif(not os.path.exists(REFERENCE_TABLES), reason="reference tree not mounted")
def test_generated_rules_equal_the_reference_tables_bit_for_bit():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kronrod_tables.py"), "--check"],
                         check=True, capture_output=True, text=True).stdout
    assert out.count("max |diff|") == 6 and out.count("nodes 0.0e+00") == 6
    assert out.count("Kronrod weights 0.0e+00") == 6 and out.count("Gauss weights 0.0e+00") == 6
