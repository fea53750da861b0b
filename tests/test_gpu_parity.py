"""GPU: parity of the CUDA path, called through the C-ABI, with the oracle:
the reference's known-answer values, the committed golden vectors produced by
the unmodified reference, and the CPU restatement run live on the same seeded
inputs.  Tolerance (BASELINE.json north_star): parameters and chi2 to 1e-10
relative, identical iteration count and accept/reject sequence."""
import numpy as np
import pytest

import backends
import scenarios
from test_oracle import _check_kat, _run_exception_sections

pytestmark = pytest.mark.gpu


def _host_threads():
    import os
    return max(1, min(os.cpu_count() or 1, 32))


@pytest.mark.parametrize("sc", scenarios.C1_SCENARIOS, ids=lambda s: s["name"])
def test_reference_known_answers(sc, c1):
    """Every KAT section of c++/tests/lm_solver.cpp, at 1e-11 relative
    (the reference asserts 1e-14 of itself; the north star allows 1e-10)."""
    res, solver = scenarios.run(sc, backends.product_solver(), c1)
    _check_kat(sc, res, c1, tol_floor=1e-11)
    assert solver.launches() > 0


@pytest.mark.parametrize("sc", scenarios.ALL_SCENARIOS, ids=lambda s: s["name"])
def test_golden_reference_runs(sc, c1, golden):
    """chi2, parameters, iteration count and the accept/reject trace of the
    unmodified reference (tests/golden/golden_ref.json)."""
    res, _ = scenarios.run(sc, backends.product_solver(), c1)
    n = scenarios.compare_to_golden(res, golden["scenarios"][sc["name"]], rtol=1e-10, label="device")
    assert n >= min(3, len(golden["scenarios"][sc["name"]]["trace"]))


def test_access_functions(c1, golden):
    """c++/tests/lm_solver.cpp:222-276: getter checksums, dof, JTJ^-1."""
    sc = scenarios.C1_SCENARIOS[0]
    s, lam, _, _ = scenarios.build(sc, backends.product_solver(), c1)
    np.testing.assert_allclose(s.getValue(c1["fix_d"][2]), 2.960644529912441, rtol=1e-13)
    s.fit(lam)
    assert s.degreesOfFreedom() == 195
    np.testing.assert_allclose(s.getValue(c1["fix_d"][2]), 36.39905496310919, rtol=1e-12)
    np.testing.assert_allclose(np.sum(s.getJacobian()), 353.6485673748526, rtol=1e-11)
    np.testing.assert_allclose(np.sum(s.getJTJ()[:5]), 580.3488115472484, rtol=1e-11)
    np.testing.assert_allclose(np.sum(s.getDTD()[:5]), 34340.67196549198, rtol=1e-11)
    np.testing.assert_allclose(np.sum(s.getLeftSide()[:5]), 614.6894835127404, rtol=1e-11)
    np.testing.assert_allclose(np.sum(s.getRightSide()[:5]), 4410.585412402701, rtol=1e-11)
    np.testing.assert_allclose(np.sum(s.getResiduals()), 213.3530475167945, rtol=1e-11)
    g = golden["scenarios"]["indexing_all"]["sums"]
    np.testing.assert_allclose(s.getJTJ(), g["jtj"], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(s.getDTD(), g["dtd"], rtol=1e-10)
    np.testing.assert_allclose(s.getRightSide(), g["right"], rtol=1e-10)
    np.testing.assert_allclose(s.getLeftSide(), g["left"], rtol=1e-10, atol=1e-9)
    prod = s.getInvJTJ().reshape(5, 5) @ s.getJTJ().reshape(5, 5)
    np.testing.assert_allclose(prod, np.eye(5), atol=1e-10)


@pytest.mark.parametrize("name,kwargs", [
    ("C3", dict(n_sets=300, n_points=200, iteration_limit=3)),   # n = 1202 columns
    ("C5", dict(n_sets=60, n_points=64, iteration_limit=2)),     # 4 shared + 3 local
    ("C2", dict(n_sets=3, n_points=5000, iteration_limit=3)),    # shared only
])
def test_block_arrow_inverse_of_jtj(name, kwargs, c1):
    """getInvJTJ (lm_solver.cpp:578-585 inverts the dense matrix) from the
    block-arrow factors, any n: times getJTJ it is the identity; its diagonal
    getter agrees; for n <= 304 it equals the oracle's dense inverse."""
    from gadfit_b200 import LMsolver, workloads as W
    w = W.WORKLOADS[name](**kwargs)
    s = LMsolver(w.model)
    w.setup(s)
    s.fit(w.lambda0)
    n = int(round(np.sqrt(s.getJTJ().size)))
    jtj = s.getJTJ().reshape(n, n)
    inv = s.getInvJTJ().reshape(n, n)
    scale = np.sqrt(np.diag(jtj))
    # conditioning-independent check: D^-1 JTJ D^-1 times D inv D
    prod = (jtj / scale[:, None] / scale[None, :]) @ (inv * scale[:, None] * scale[None, :])
    np.testing.assert_allclose(prod, np.eye(n), atol=1e-7)
    np.testing.assert_allclose(inv, inv.T, rtol=1e-9, atol=1e-30)
    np.testing.assert_allclose(s.getInvJTJDiagonal(), np.diag(inv), rtol=1e-12)
    if n <= 304:
        o = backends.oracle_solver(1)(w.model)
        w.setup(o)
        o.fit(w.lambda0)
        np.testing.assert_allclose(inv, o.getInvJTJ().reshape(n, n), rtol=1e-6, atol=1e-12 * np.abs(inv).max())


def test_exceptions(c1):
    """c++/tests/lm_solver.cpp:278-365 -> status codes -> same exception names"""
    import gadfit_b200 as g
    _run_exception_sections(backends.product_solver(), c1, g)


def test_warm_restart_continues_the_fit(c1):
    """Parameters persist in the solver: a second fit() continues
    (reference c++/tests/numerical_integration.cpp:63-72 relies on this)."""
    sc = scenarios.C1_SCENARIOS[0]
    a, lam, _, _ = scenarios.build(sc, backends.product_solver(), c1)
    a.settings.iteration_limit = 2
    a.fit(lam)
    b, _, _, _ = scenarios.build(sc, backends.oracle_solver(1), c1)
    b.settings.iteration_limit = 2
    b.fit(lam)
    a.fit(0.01)
    b.fit(0.01)
    np.testing.assert_allclose(a.chi2(), b.chi2(), rtol=1e-11)
    np.testing.assert_allclose(a.getParValue(1), b.getParValue(1), rtol=1e-11)


@pytest.mark.parametrize("name,kwargs", [
    ("C2", dict(n_sets=5, n_points=20_000)),
    ("C3", dict(n_sets=300, n_points=700)),
    ("C4", dict(n_sets=12, n_points=1500)),
    ("C5", dict(n_sets=250, n_points=96)),
])
def test_against_block_arrow_oracle_at_sizes_the_reference_cannot_run(name, kwargs, c1):
    """Sizes beyond the dense reference (n up to 1202 columns): the CUDA path
    vs the CPU restatement (block-arrow mode) on the same seeded inputs."""
    sc = dict(name=name, kind="workload", workload=name, kwargs=kwargs)
    got, gs = scenarios.run(sc, backends.product_solver(), c1)
    want, _ = scenarios.run(sc, backends.oracle_solver(2), c1, n_threads=8)
    n = scenarios.compare_to_golden(got, want, rtol=1e-10, label="device-vs-oracle")
    assert n >= 3
    assert gs.launches() > 0


@pytest.mark.parametrize("name,kwargs", [
    ("C2", dict(n_sets=2, n_points=300_000, iteration_limit=3)),
    ("C3", dict(n_sets=40, n_points=30_000, iteration_limit=3)),
])
def test_many_tiles_per_cta_against_the_oracle(name, kwargs, c1):
    """Data sets long enough that every CTA of the Jacobian sweep walks several
    tiles (both staging buffers of the warp-specialised kernel in use, data-set
    changes in the middle of a CTA's range)."""
    sc = dict(name=f"{name}_long", kind="workload", workload=name, kwargs=kwargs)
    got, gs = scenarios.run(sc, backends.product_solver(), c1)
    want, _ = scenarios.run(sc, backends.oracle_solver(2), c1, n_threads=_host_threads())
    n = scenarios.compare_to_golden(got, want, rtol=1e-10, label="device-vs-oracle")
    assert n >= 3


def test_full_size_c3_invariants(c1):
    """BASELINE config 3 at FULL size (1000 x 4096, n = 4002), which the
    reference cannot run: size-independent properties.
    (1) chi2 of the chi2 sweep == weighted residual norm of the Jacobian sweep
        (two independent kernels, linear loss) -- a checksum of checksums;
    (2) reversing the order of the data sets leaves the shared parameters
        unchanged (block-arrow structure);
    (3) chi2 decreases monotonically over the accepted steps."""
    from gadfit_b200 import LMsolver, workloads as W
    w = W.c3(iteration_limit=6)
    s = LMsolver(w.model)
    w.setup(s)
    s.fit(w.lambda0)
    tr = s.trace()
    acc = [t for t in tr if t["decision"] == 0]
    assert len(acc) == s.iterations() == 6
    assert all(t["new_chi2"] < t["old_chi2"] for t in acc)
    res = s.getResiduals()
    # residuals belong to the last Jacobian evaluation = old_chi2 of the last
    # accepted trial chain's first entry
    last_iter = tr[-1]["iteration"]
    first_of_last = [t for t in tr if t["iteration"] == last_iter][0]
    np.testing.assert_allclose(np.sum(res * res), first_of_last["old_chi2"], rtol=1e-12)
    shared = [s.getParValue(k) for k in w.shared]
    # same problem, data sets in reverse order
    r = LMsolver(w.model)
    for i in reversed(range(w.n_sets)):
        xs, ys, es = w.slice(i)
        r.addDataset(xs, ys, es)
    for k in w.shared:
        r.setPar(k, float(w.start[0, k]), True)
    for j, i in enumerate(reversed(range(w.n_sets))):
        for k in w.local:
            r.setPar(k, float(w.start[i, k]), True, j)
    r.settings.iteration_limit = 6
    r.fit(w.lambda0)
    np.testing.assert_allclose([r.getParValue(k) for k in w.shared], shared, rtol=1e-10)
    np.testing.assert_allclose(r.getParValue(0, 0), s.getParValue(0, w.n_sets - 1), rtol=1e-10)


def test_splitting_a_data_set_does_not_change_a_shared_fit(c1):
    """All parameters shared: fitting one data set of 2n points equals fitting
    its two halves as separate data sets (the normal equations are sums)."""
    from gadfit_b200 import LMsolver, workloads as W
    w = W.c2(n_sets=2, n_points=30_000, iteration_limit=5)
    a = LMsolver(w.model)
    w.setup(a)
    a.fit(w.lambda0)
    b = LMsolver(w.model)
    for i in range(2):
        xs, ys, es = w.slice(i)
        h = 12_345
        b.addDataset(xs[:h], ys[:h], es[:h])
        b.addDataset(xs[h:], ys[h:], es[h:])
    for k in w.shared:
        b.setPar(k, float(w.start[0, k]), True)
    b.settings.iteration_limit = 5
    b.fit(w.lambda0)
    np.testing.assert_allclose(b.chi2(), a.chi2(), rtol=1e-11)
    np.testing.assert_allclose([b.getParValue(k) for k in range(20)],
                               [a.getParValue(k) for k in range(20)], rtol=1e-9)


def test_empty_and_single_point_data_sets(c1):
    """Edge cases of the tile bookkeeping: an empty data set and a 1-point one."""
    from gadfit_b200 import LMsolver
    d = c1["fix_d"]
    outs = []
    for make in (backends.product_solver(), backends.oracle_solver(1)):
        s = make("exponential")
        s.addDataset(c1["x_data_1"], c1["y_data_1"])
        s.addDataset(np.zeros(0), np.zeros(0))
        s.addDataset(c1["x_data_2"][:1].copy(), c1["y_data_2"][:1].copy())
        s.setPar(0, d[0], True, 0)
        s.setPar(2, d[1], True, 0)
        s.setPar(0, d[4], False, 1)
        s.setPar(2, d[5], False, 1)
        s.setPar(0, d[4], False, 2)
        s.setPar(2, d[5], False, 2)
        s.setPar(1, d[3], True)
        s.settings.iteration_limit = 3
        s.fit(1.0)
        outs.append([s.chi2(), s.getParValue(1), s.getParValue(0, 0), s.getParValue(2, 0)])
    np.testing.assert_allclose(outs[0], outs[1], rtol=1e-10)


def test_fp64_peak_microbenchmark_is_sane():
    from gadfit_b200 import measure_fp64_peak
    peak = measure_fp64_peak()
    assert 5e12 < peak < 80e12


def _consistency_checks(s, n_iterations):
    """Size-independent properties of a finished fit (used at BASELINE full
    sizes where no CPU oracle finishes in test time): chi2 decreases over the
    accepted steps; lambda follows the reference schedule exactly
    (lm_solver.cpp:463,471); the chi2 the solver reports equals an independent
    chi2 sweep at the final parameters."""
    tr = s.trace()
    acc = [t for t in tr if t["decision"] == 0]
    assert len(acc) == s.iterations() == n_iterations
    assert all(t["new_chi2"] < t["old_chi2"] for t in acc)
    for a, b in zip(tr, tr[1:]):
        want = a["lambda_"] / 10.0 if a["decision"] == 0 else a["lambda_"] * 10.0
        assert b["lambda_"] == want
    np.testing.assert_allclose(s.chi2(), acc[-1]["new_chi2"], rtol=1e-13)


@pytest.mark.parametrize("name,kwargs,what", [
    # BASELINE config 3 at FULL size, BASELINE's own settings (10 iterations):
    # n = 4002 columns, which the dense reference cannot run
    ("C3", dict(iteration_limit=10), "full size 1000 x 4096"),
    # BASELINE config 4, full-size data sets (65 536 points, geodesic
    # acceleration, 8 iterations), a quarter of the data sets: n = 131 columns
    ("C4", dict(n_sets=64, n_points=65536, iteration_limit=8), "64 x 65536"),
    # BASELINE config 5, full-size data sets, 2000 of them: n = 6004 columns
    ("C5", dict(n_sets=2000, n_points=256, iteration_limit=6), "2000 x 256"),
], ids=["C3_full", "C4_64x65536", "C5_2000x256"])
def test_baseline_shapes_against_the_block_arrow_oracle(name, kwargs, what, c1):
    """The BASELINE shapes at (or near) full size against the CPU restatement
    in block-arrow mode on all host threads (seconds to a minute): identical
    accept/reject sequence and lambda, chi2 and EVERY parameter of every data
    set to 1e-10 -- not invariants."""
    sc = dict(name=f"{name}_{what.replace(' ', '')}", kind="workload", workload=name, kwargs=kwargs)
    got, gs = scenarios.run(sc, backends.product_solver(), c1)
    want, _ = scenarios.run(sc, backends.oracle_solver(2), c1, n_threads=_host_threads())
    n = scenarios.compare_to_golden(got, want, rtol=1e-10, label="device-vs-oracle")
    assert n >= 3
    assert n.record["trials_identical"] == len(want["trace"]), n.record
    assert gs.launches() > 0


def test_full_size_c4_single_integral_with_acceleration(c1):
    """BASELINE config 4 at FULL size (256 x 65 536 points, fixed 15-node
    Kronrod panel per point, geodesic acceleration), which no CPU oracle
    finishes in test time: invariants, and the first 64 data sets' problem
    is the one test_baseline_shapes_against_the_block_arrow_oracle pins to
    1e-10 on every parameter."""
    from gadfit_b200 import LMsolver, workloads as W
    w = W.c4(iteration_limit=4)
    s = LMsolver(w.model)
    w.setup(s)
    s.fit(w.lambda0)
    _consistency_checks(s, 4)
    assert s.timings()["omega"][1] >= 4  # the omega sweep ran in every iteration


def test_full_size_c5_double_integral_block_arrow(c1):
    """BASELINE config 5 at FULL size (100 000 data sets x 256 points, 15 x 15
    Kronrod per point, n = 300 004 columns): the Schur-solve stress test.
    Invariants + invariance of the shared parameters under reversing the order
    of the data sets."""
    from gadfit_b200 import LMsolver, workloads as W
    w = W.c5(iteration_limit=3)
    s = LMsolver(w.model)
    w.setup(s)
    s.fit(w.lambda0)
    _consistency_checks(s, 3)
    shared = [s.getParValue(k) for k in w.shared]
    r = LMsolver(w.model)
    order = list(reversed(range(w.n_sets)))
    for i in order:
        xs, ys, es = w.slice(i)
        r.addDataset(xs, ys, es)
    for k in w.shared:
        r.setPar(k, float(w.start[0, k]), True)
    for j, i in enumerate(order):
        for k in w.local:
            r.setPar(k, float(w.start[i, k]), True, j)
    r.settings.iteration_limit = 3
    r.fit(w.lambda0)
    np.testing.assert_allclose([r.getParValue(k) for k in w.shared], shared, rtol=1e-10)
    np.testing.assert_allclose(r.getParValue(0, 0), s.getParValue(0, w.n_sets - 1), rtol=1e-10)
    assert s.degreesOfFreedom() == w.n_total - (4 + 3 * w.n_sets)


def test_step_vectors_and_omega_with_acceleration(c1):
    """delta1, delta2 and omega of the last trial step (lm_solver.cpp:348-399)
    against the oracle, geodesic acceleration on."""
    sc = [s for s in scenarios.C1_SCENARIOS if s["name"] == "geodesic_acceleration"][0]
    a, lam, _, _ = scenarios.build(sc, backends.product_solver(), c1)
    b, _, _, _ = scenarios.build(sc, backends.oracle_solver(1), c1)
    a.fit(lam)
    b.fit(lam)
    np.testing.assert_allclose(a.getDelta1(), b.getDelta1(), rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(a.getDelta2(), b.getDelta2(), rtol=1e-8, atol=1e-14)
    np.testing.assert_allclose(a.getOmega(), b._array(9), rtol=1e-9, atol=1e-16)


def _is_accelerated(sc):
    """Geodesic acceleration (an omega sweep in the middle of the solve) keeps
    the stage-wise path."""
    return (sc.get("workload") == "C4"
            or (sc.get("settings") or {}).get("acceleration_threshold", -1.0) > 0.0)


@pytest.mark.parametrize(
    "sc",
    [s for s in scenarios.ALL_SCENARIOS if not _is_accelerated(s)]
    + [dict(name="C3_small", kind="workload", workload="C3", kwargs=dict(n_sets=300, n_points=700)),
       dict(name="C5_small", kind="workload", workload="C5", kwargs=dict(n_sets=250, n_points=96)),
       dict(name="C2_small", kind="workload", workload="C2", kwargs=dict(n_sets=5, n_points=20_000)),
       # 100 data sets: a partly filled CTA; 1500 and 2048: more data sets than
       # the cluster has set-threads (1024), i.e. two passes (2048 = the limit)
       dict(name="C3_100", kind="workload", workload="C3", kwargs=dict(n_sets=100, n_points=200)),
       dict(name="C3_1500", kind="workload", workload="C3", kwargs=dict(n_sets=1500, n_points=64)),
       dict(name="C5_2048", kind="workload", workload="C5", kwargs=dict(n_sets=2048, n_points=24))],
    ids=lambda s: s["name"])
def test_fused_small_problem_path_is_bit_identical_to_the_stagewise_path(sc, c1, monkeypatch):
    """The latency-bound regime (one GPU, <= 2048 data sets, no acceleration)
    runs a trial step as 4 launches (k_solve_small / k_finish_trial) instead of
    16.  Same per-data-set code, same summation order: every trial record and
    every parameter must be identical to the last bit, with fewer launches."""
    monkeypatch.setenv("GADFIT_B200_FUSED", "1")
    fused, fs = scenarios.run(sc, backends.product_solver(), c1)
    monkeypatch.setenv("GADFIT_B200_FUSED", "0")
    staged, ss = scenarios.run(sc, backends.product_solver(), c1)
    assert fused["trace"] == staged["trace"]
    assert fused["chi2"] == staged["chi2"]
    assert fused["pars"] == staged["pars"]
    assert fused["iterations"] == staged["iterations"]
    assert fs.launches() > 0
    if sc["name"] != "no_iteration_limit":
        # (without an iteration limit the device-driven loop pre-enqueues a batch
        # of trial slots; those past the end of the fit skip themselves but count)
        assert fused["iterations"] == 0 or fs.launches() < ss.launches()
    for getter in ("getJTJ", "getDTD", "getRightSide", "getLeftSide"):
        if fused["dof"] < 10_000:  # dense getters: small cases only
            np.testing.assert_array_equal(getattr(fs, getter)(), getattr(ss, getter)())


@pytest.mark.parametrize(
    "sc",
    [s for s in scenarios.ALL_SCENARIOS if not _is_accelerated(s)]
    + [dict(name="C3_small", kind="workload", workload="C3", kwargs=dict(n_sets=300, n_points=700)),
       dict(name="C2_small", kind="workload", workload="C2", kwargs=dict(n_sets=5, n_points=20_000))],
    ids=lambda s: s["name"])
def test_device_driven_loop_is_bit_identical_to_the_host_driven_loop(sc, c1, monkeypatch):
    """Latency-bound regime: with the accept / reject decision, the revert and
    the trace on the device (the kernels of a trial slot skip themselves by
    LmState::phase; the host reads the state once per batch of slots) every
    trial record and every parameter is identical to the last bit to the loop
    that reads one record per trial."""
    monkeypatch.setenv("GADFIT_B200_DEVICE_LOOP", "1")
    dev, ds = scenarios.run(sc, backends.product_solver(), c1)
    monkeypatch.setenv("GADFIT_B200_DEVICE_LOOP", "0")
    host, hs = scenarios.run(sc, backends.product_solver(), c1)
    assert dev["trace"] == host["trace"]
    assert dev["chi2"] == host["chi2"]
    assert dev["pars"] == host["pars"]
    assert dev["iterations"] == host["iterations"]
    assert ds.stopReason() == hs.stopReason()
    assert ds.timings()["jacobian"][1] == hs.timings()["jacobian"][1]
    assert ds.timings()["chi2"][1] == hs.timings()["chi2"][1]


def test_distinct_solvers_on_concurrent_host_threads(c1):
    """SURVEY 8(b) threading row (reference doc/user_guide.tex:778): one solver
    is not re-entrant, but distinct solvers may run concurrently from different
    host threads.  Eight threads, each fitting its own known-answer scenario on
    its own handle (own stream), must all reproduce the reference's values."""
    import threading
    todo = (scenarios.C1_SCENARIOS * 2)[:8]
    make = backends.product_solver()
    results, errors = [None] * len(todo), []

    def work(i, sc):
        try:
            for _ in range(3):
                results[i], _s = scenarios.run(sc, make, c1)
        except Exception as e:  # noqa: BLE001
            errors.append((sc["name"], repr(e)))

    threads = [threading.Thread(target=work, args=(i, sc)) for i, sc in enumerate(todo)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for sc, res in zip(todo, results):
        _check_kat(sc, res, c1, tol_floor=1e-11)


# ---- round 2: kernel variants chosen by shape must agree with each other -----

@pytest.mark.parametrize("name,kwargs,variants", [
    ("C3", dict(n_sets=40, n_points=3000, iteration_limit=3), ("reg", "sp", "ws")),
    ("C2", dict(n_sets=3, n_points=40_000, iteration_limit=3), ("ws", "sp")),
    ("C4", dict(n_sets=6, n_points=2000, iteration_limit=3), ("sp", "ws")),
])
def test_jacobian_sweep_variants_agree(name, kwargs, variants, monkeypatch):
    """The three Jacobian sweeps (warp-specialised DMMA, software-pipelined
    DMMA, accumulated in registers; csrc/sweep_kernels.cuh) differ only in the
    order of the sums: J^T J, J^T r to 1e-12 of the diagonal scale, identical
    decisions, parameters to 1e-10."""
    from gadfit_b200 import LMsolver, workloads as W
    results = []
    for v in variants:
        monkeypatch.setenv("GADFIT_B200_JAC", v)
        w = W.WORKLOADS[name](**kwargs)
        s = LMsolver(w.model)
        w.setup(s)
        s.fit(w.lambda0)
        n = int(round(np.sqrt(s.getJTJ().size)))
        results.append(dict(jtj=s.getJTJ().reshape(n, n), rhs=s.getRightSide(),
                            decisions=[t["decision"] for t in s.trace()], chi2=s.chi2(),
                            pars=np.array([[s.getParValue(k, j) for k in range(w.n_pars)]
                                           for j in range(w.n_sets)])))
    ref = results[0]
    scale = np.sqrt(np.abs(np.diag(ref["jtj"])))
    for r in results[1:]:
        assert r["decisions"] == ref["decisions"]
        np.testing.assert_allclose(r["jtj"] / scale[:, None] / scale[None, :],
                                   ref["jtj"] / scale[:, None] / scale[None, :], rtol=0, atol=1e-12)
        np.testing.assert_allclose(r["rhs"] / scale, ref["rhs"] / scale, rtol=0,
                                   atol=1e-11 * np.max(np.abs(ref["rhs"] / scale)))
        np.testing.assert_allclose(r["chi2"], ref["chi2"], rtol=1e-12)
        np.testing.assert_allclose(r["pars"], ref["pars"], rtol=1e-10, atol=1e-13)


def test_staged_chi2_sweep_is_bit_identical_to_the_direct_one(monkeypatch):
    """The chi2 sweep with inputs staged through shared memory (long tile
    ranges) adds the same terms in the same order as the one with direct loads:
    the whole trace must be identical to the last bit."""
    from gadfit_b200 import LMsolver, workloads as W
    traces = []
    for v in ("direct", "staged"):
        monkeypatch.setenv("GADFIT_B200_CHI2", v)
        w = W.WORKLOADS["C2"](n_sets=3, n_points=70_001, iteration_limit=3)  # ragged last tile
        s = LMsolver(w.model)
        w.setup(s)
        s.fit(w.lambda0)
        traces.append(([(t["decision"], t["new_chi2"], t["lambda_"]) for t in s.trace()], s.chi2()))
    assert traces[0] == traces[1]


def test_phase_timers_can_be_switched_off(c1):
    """gadfit_b200_set_phase_timers(0): no event records on the stream, the fit
    itself is unchanged; the whole-fit timer keeps running."""
    sc = scenarios.C1_SCENARIOS[0]
    a, lam, _, _ = scenarios.build(sc, backends.product_solver(), c1)
    a.fit(lam)
    b, _, _, _ = scenarios.build(sc, backends.product_solver(), c1)
    b.setPhaseTimers(False)
    b.fit(lam)
    assert [t["new_chi2"] for t in a.trace()] == [t["new_chi2"] for t in b.trace()]
    ta, tb = a.timings(), b.timings()
    assert ta["jacobian"][0] > 0.0 and tb["jacobian"][0] == 0.0
    assert tb["main"][0] > 0.0 and tb["jacobian"][1] == ta["jacobian"][1]
