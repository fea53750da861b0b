"""Library handles for the tests: the product (CUDA) and the two oracle
libraries (test infrastructure; see oracle/)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from gadfit_b200._capi import Api
from gadfit_b200.lm_solver import LMsolver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libgadfit_ref.so")
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "libgadfit_oracle.so")

_cache: dict = {}


def _env():
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    return env


def oracle_api() -> Api:
    """The CPU restatement; compiled on demand (g++, a few seconds)."""
    if "oracle" not in _cache:
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"],
                       check=True, env=_env(), stdout=subprocess.DEVNULL)
        api = Api(ORACLE_LIB, "gadfit_oracle_", product=False)
        api.lib.gadfit_oracle_set_mode.argtypes = [C.c_void_p, C.c_int]
        _cache["oracle"] = api
    return _cache["oracle"]


def has_ref() -> bool:
    return os.path.exists(REF_LIB)


def ref_api() -> Api:
    """The unmodified reference behind oracle/ref_harness.cpp (prebuilt by
    __graft_entry__.build() where /root/reference is mounted)."""
    if "ref" not in _cache:
        _cache["ref"] = Api(REF_LIB, "gadfit_ref_", product=False)
    return _cache["ref"]


def oracle_solver(mode: int = 0):
    """mode: 0 automatic, 1 dense (reference's loops), 2 block-arrow Schur."""
    api = oracle_api()

    def make(model: str) -> LMsolver:
        s = LMsolver(model, _api=api)
        api.lib.gadfit_oracle_set_mode(s._h, mode)
        return s
    return make


def ref_solver():
    api = ref_api()
    return lambda model: LMsolver(model, _api=api)


def product_solver():
    """The CUDA library; the functions of the reference's nested / direct
    double-integral tests come from the test-model plug-in."""
    import gadfit_b200
    gadfit_b200.load_test_models()
    return lambda model: LMsolver(model)
