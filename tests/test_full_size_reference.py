"""GPU, slow: BASELINE config 2 at FULL size (64 data sets x 1M points, 20
shared parameters) fitted by the CUDA library and by the UNMODIFIED reference
(oracle/_ref, all host threads) on identical inputs -- the reference's
optimised configuration (lapack.cpp + blocked OpenBLAS dsyrk) by default, and
its default in-tree linear algebra (lapack_fallback.cpp, sequential dsyrk) as
a second comparand with GADFIT_SLOW=1.  Plus the rounding probe: the same
(smaller) instance a third time with the normal equations accumulated in long
double, to say which double-precision solver is the noisy one.

Iterations: GADFIT_FULL_ITERS (default 2: ~20 s per reference iteration on the
GPU box's 16 host threads; BASELINE's setting is 10).  A JSON record with the
per-parameter differences goes to $GADFIT_PARITY_DIR (default gpurun_out/).
"""
import ctypes as C
import json
import os
import time

import numpy as np
import pytest

import backends

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
pytestmark = [pytest.mark.gpu, pytest.mark.slow]

ITERS = int(os.environ.get("GADFIT_FULL_ITERS", "2"))
SLOW = os.environ.get("GADFIT_SLOW", "0") == "1"
COMPARANDS = [("blas", "libgadfit_ref_blas.so", "lapack.cpp + OpenBLAS (blocked dsyrk)")]
if SLOW:
    COMPARANDS.append(("in_tree", "libgadfit_ref_fast.so", "lapack_fallback.cpp (sequential dsyrk)"))


def _record(name, rec):
    out = os.environ.get("GADFIT_PARITY_DIR", os.path.join(ROOT, "gpurun_out"))
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, name), "w") as f:
            json.dump(rec, f, indent=1)
    except OSError:
        pass


def _fit(w, api=None, mode=None):
    from gadfit_b200 import LMsolver
    s = LMsolver(w.model, _api=api)
    if mode is not None:
        api.lib.gadfit_oracle_set_mode.argtypes = [C.c_void_p, C.c_int]
        api.lib.gadfit_oracle_set_mode(s._h, mode)
    w.setup(s)
    s.settings.n_threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    s.fit(w.lambda0)
    dt = time.perf_counter() - t0
    pars = np.array([s.getParValue(k) for k in range(w.n_pars)])
    return pars, s.trace(), s.iterations(), dt


@pytest.mark.parametrize("tag,lib,what", COMPARANDS, ids=[c[0] for c in COMPARANDS])
def test_full_size_c2_against_the_unmodified_reference(tag, lib, what):
    from gadfit_b200 import workloads as W
    from gadfit_b200._capi import Api
    path = os.path.join(REF_DIR, lib)
    if not os.path.exists(path):
        pytest.skip(f"{lib} not built (needs /root/reference at build time)")
    w = W.c2(iteration_limit=ITERS)
    gp, gtr, git, t_gpu = _fit(w)
    rp, rtr, rit, t_ref = _fit(w, Api(path, "gadfit_ref_", product=False))
    rel = np.abs(gp - rp) / np.abs(rp)
    chi2_rel = max(abs(a["new_chi2"] - b["new_chi2"]) / abs(b["new_chi2"]) for a, b in zip(gtr, rtr))
    rec = dict(
        workload=f"C2 full size: {w.n_sets} x {w.n_points} points, 20 shared parameters, {ITERS} iterations",
        reference=f"unmodified reference, {what}, {os.cpu_count()} host threads",
        seconds=dict(gpu_fit_incl_upload=t_gpu, reference_fit=t_ref),
        iterations=dict(gpu=git, reference=rit),
        decisions=dict(gpu=[t["decision"] for t in gtr], reference=[t["decision"] for t in rtr]),
        lambdas_identical=[t["lambda_"] for t in gtr] == [t["lambda_"] for t in rtr],
        chi2_trace_max_rel_diff=chi2_rel,
        parameter_rel_diff=rel.tolist(),
        parameters_within_1e10=[bool(r <= 1e-10) for r in rel],
        parameter_max_rel_diff=float(rel.max()),
        parameters=dict(gpu=gp.tolist(), reference=rp.tolist()),
    )
    _record(f"full_size_parity_c2_{tag}.json", rec)
    print(json.dumps({k: rec[k] for k in ("seconds", "iterations", "lambdas_identical",
                                          "chi2_trace_max_rel_diff", "parameter_max_rel_diff",
                                          "parameters_within_1e10")}))
    assert git == rit == ITERS
    assert rec["decisions"]["gpu"] == rec["decisions"]["reference"]
    assert rec["lambdas_identical"]
    assert chi2_rel <= 1e-10
    # north star: 1e-10 relative on the fitted parameters
    assert rel.max() <= 1e-10, rec["parameter_rel_diff"]


def test_rounding_probe_c2():
    """C2 with 16 x 250k points three ways: CUDA library, unmodified reference
    (double), oracle restatement accumulating J^T J, J^T r and chi2 in long
    double.  The CUDA library must be at least as close to the extended
    solution as the reference is, and within 1e-10 of it."""
    from gadfit_b200 import workloads as W
    from gadfit_b200._capi import Api
    path = os.path.join(REF_DIR, "libgadfit_ref_fast.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built")
    w = W.c2(n_sets=16, n_points=250_000, iteration_limit=10)
    gpu, gtr, _, _ = _fit(w)
    ref, rtr, _, _ = _fit(w, Api(path, "gadfit_ref_", product=False))
    ext, etr, _, _ = _fit(w, backends.oracle_api(), mode=1 | 4)
    rel = lambda a, b: np.abs(a - b) / np.abs(b)  # noqa: E731
    rec = dict(
        workload=f"C2: {w.n_sets} x {w.n_points} points, 10 iterations",
        decisions_identical=[t["decision"] for t in gtr] == [t["decision"] for t in rtr]
        == [t["decision"] for t in etr],
        gpu_vs_extended=dict(max=float(rel(gpu, ext).max()), per_parameter=rel(gpu, ext).tolist()),
        reference_vs_extended=dict(max=float(rel(ref, ext).max()), per_parameter=rel(ref, ext).tolist()),
        gpu_vs_reference=dict(max=float(rel(gpu, ref).max()), per_parameter=rel(gpu, ref).tolist()),
    )
    _record("rounding_probe_c2.json", rec)
    print(json.dumps({k: (v if not isinstance(v, dict) else v["max"]) for k, v in rec.items()}))
    assert rec["decisions_identical"]
    assert rec["gpu_vs_extended"]["max"] <= 1e-10
    assert rec["gpu_vs_extended"]["max"] <= max(rec["reference_vs_extended"]["max"], 1e-12)
