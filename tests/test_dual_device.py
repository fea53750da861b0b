"""GPU: the dual-number rules of include/gadfit_b200/dual.h executed IN A
KERNEL (tests/device_dual_check.cu; CUDA's math library and FMA contraction,
as in the product kernels), all 28 elemental operations of the reference
(automatic_differentiation.cpp:59-680): value, Dir2 first and second
directional derivative, SDual gradient.  Compared with
  * the reference's known answers, c++/tests/ad_forward_mode.cpp:16-256 and
    c++/tests/ad_reverse_mode.cpp (@1e-13, the reference test's tolerance),
  * gadfit::AdVar itself, one operation at a time
    (tests/golden/dual_ops_ref.json, @1e-13),
  * the host build of the same header (@1e-13).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_dual_host import FORWARD_KAT, GRADIENT_KAT, SINGLE_OPS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _lib():
    from gadfit_b200 import build
    return C.CDLL(build.build_device_checks())


def _ulps(got, want):
    spacing = np.spacing(np.abs(want))
    return np.abs(got - want) / spacing


def test_branch_free_exp_and_reciprocal_on_the_device():
    """math_exp / math_rcp (dual.h) in a kernel against CUDA's exp() and
    1.0 / x and against numpy: <= 1 ulp of CUDA's correctly rounded division,
    <= 2 ulp between the two exp implementations, special values identical."""
    lib = _lib()
    rng = np.random.default_rng(3)
    special = np.array([0.0, -0.0, 1.0, -1.0, 709.0, 709.78, 709.79, 710.0, 1e300, -745.0, -745.2,
                        -746.0, -1e300, np.inf, -np.inf, np.nan, 1e-320, -708.4, -720.0, 2.0 ** 1023,
                        2.0 ** -1022, 3.0, 1e-300, 1e300, -7.0, 0.1])
    x = np.concatenate([special, rng.uniform(-30, 30, 200_000), rng.uniform(-700, 700, 200_000),
                        rng.uniform(-1, 1, 100_000) * 10.0 ** rng.uniform(-200, 200, 100_000)])
    n = x.size
    out = np.empty(4 * n)
    rc = lib.device_math_check(x.ctypes.data_as(C.POINTER(C.c_double)), n,
                               out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    my_exp, cuda_exp, my_rcp, cuda_rcp = out.reshape(4, n)
    with np.errstate(all="ignore"):
        np_exp = np.exp(x)
    # special values
    k = special.size
    for a, b in ((my_exp[:k], np_exp[:k]),):
        assert np.array_equal(np.isnan(a), np.isnan(b))
        fin = np.isfinite(b) & (b > 1e-300)
        assert np.array_equal(a[~fin & ~np.isnan(b)], b[~fin & ~np.isnan(b)]) or np.allclose(
            a[~fin & ~np.isnan(b)], b[~fin & ~np.isnan(b)], rtol=0, atol=1e-323)
    normal = np.isfinite(np_exp) & (np_exp > 1e-300) & (np_exp < 1e300)
    assert _ulps(my_exp[normal], np_exp[normal]).max() <= 1.5
    assert _ulps(my_exp[normal], cuda_exp[normal]).max() <= 2.0
    # reciprocal: normal arguments whose reciprocal is normal
    ok = np.isfinite(x) & (np.abs(x) > 1e-300) & (np.abs(x) < 1e300)
    assert _ulps(my_rcp[ok], cuda_rcp[ok]).max() <= 1.0
    assert my_rcp[0] == np.inf and my_rcp[1] == -np.inf           # 1 / +-0
    assert my_rcp[13] == 0.0 and my_rcp[14] == 0.0 and np.isnan(my_rcp[15])
    print("exp: %.2f ulp vs numpy, %.2f vs CUDA; reciprocal: %.2f ulp vs CUDA division" % (
        _ulps(my_exp[normal], np_exp[normal]).max(), _ulps(my_exp[normal], cuda_exp[normal]).max(),
        _ulps(my_rcp[ok], cuda_rcp[ok]).max()))


@pytest.fixture(scope="module")
def device_values():
    lib = _lib()
    lib.device_dual_check.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int),
                                      C.POINTER(C.c_int)]
    lib.device_dual_check_name.restype = C.c_char_p
    out = (C.c_double * 4096)()
    n, w = C.c_int(), C.c_int()
    rc = lib.device_dual_check(out, 4096, C.byref(n), C.byref(w))
    assert rc == 0, f"device_dual_check failed: {rc}"
    vals = np.array(out[: n.value * w.value]).reshape(n.value, w.value)
    return {lib.device_dual_check_name(i).decode(): vals[i] for i in range(n.value)}


@pytest.fixture(scope="module")
def host_values(tmp_path_factory):
    exe = tmp_path_factory.mktemp("dual") / "host_dual_check"
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.run(["g++", "-std=c++20", "-O2", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "host_dual_check.cpp"), "-o", str(exe)],
                   check=True, env=env)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    return {ln.split()[0]: np.array([float(v) for v in ln.split()[1:]]) for ln in out.splitlines()}


@pytest.mark.parametrize("name", sorted(FORWARD_KAT))
def test_device_forward_mode_known_answers(name, device_values):
    np.testing.assert_allclose(device_values[name][:3], FORWARD_KAT[name], rtol=1e-13, atol=0.0)


@pytest.mark.parametrize("name", sorted(GRADIENT_KAT))
def test_device_static_sparsity_gradients(name, device_values):
    np.testing.assert_allclose(device_values[name], GRADIENT_KAT[name], rtol=1e-13, atol=0.0)


@pytest.mark.parametrize("name", sorted(SINGLE_OPS))
def test_device_single_operations_against_the_reference_advar(name, device_values):
    np.testing.assert_allclose(device_values[name][:3], SINGLE_OPS[name], rtol=1e-13, atol=1e-300)


def test_device_matches_the_host_build_of_the_same_header(device_values, host_values):
    assert sorted(device_values) == sorted(host_values)
    assert len(device_values) == 23 + 2 * 25
    worst = 0.0
    for name, dev in device_values.items():
        host = host_values[name]
        np.testing.assert_allclose(dev, host, rtol=1e-13, atol=1e-300, err_msg=name)
        worst = max(worst, float(np.max(np.abs(dev - host) / np.maximum(np.abs(host), 1e-300))))
    print(f"device vs host dual numbers: worst relative difference {worst:.2e}")
