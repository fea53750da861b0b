"""CPU: pin the oracle.  The restatement (oracle/oracle_solver.cpp) and --
where it was built -- the compiled reference (oracle/_ref) are checked against
the known-answer values of the reference's own tests
(c++/tests/lm_solver.cpp) and against the committed golden vectors."""
import numpy as np
import pytest

import backends
import scenarios


def _check_kat(sc, res, c1, tol_floor=0.0):
    d = c1["fix_d"]
    chi2, tau, i00, bgr0, i01, bgr1 = sc["kat"]
    tol = max(sc["tol"], tol_floor)
    i0a, i0b = sc["i0"]
    got = [res["chi2"], res["pars"][0][1], res["pars"][0][0], res["pars"][0][2],
           res["pars"][1][0], res["pars"][1][2]]
    want = [chi2, tau, i00 if i00 is not None else d[i0a], bgr0 if bgr0 is not None else d[1],
            i01 if i01 is not None else d[i0b], bgr1 if bgr1 is not None else d[5]]
    for g, w in zip(got, want):
        if w is not None:
            np.testing.assert_allclose(g, w, rtol=tol, atol=0.0,
                                       err_msg=f"{sc['name']} (lm_solver.cpp:{sc['lines']})")
    # the shared parameter is the same in both data sets
    assert res["pars"][0][1] == res["pars"][1][1]


@pytest.mark.parametrize("sc", scenarios.C1_SCENARIOS, ids=lambda s: s["name"])
@pytest.mark.parametrize("mode", [1, 2], ids=["dense", "block_arrow"])
def test_oracle_reproduces_reference_known_answers(sc, mode, c1):
    """KATs of c++/tests/lm_solver.cpp at the tolerance that file uses (1e-14
    relative; 1e-8 for the free-running case, 1e-12 for dof = 0), relaxed to
    4e-14 because the restatement differentiates in forward mode."""
    res, _ = scenarios.run(sc, backends.oracle_solver(mode), c1)
    _check_kat(sc, res, c1, tol_floor=4e-14)


@pytest.mark.parametrize("sc", scenarios.ALL_SCENARIOS, ids=lambda s: s["name"])
@pytest.mark.parametrize("mode", [1, 2], ids=["dense", "block_arrow"])
def test_oracle_matches_golden_reference_runs(sc, mode, c1, golden):
    res, _ = scenarios.run(sc, backends.oracle_solver(mode), c1)
    n = scenarios.compare_to_golden(res, golden["scenarios"][sc["name"]], rtol=1e-11,
                                    label="oracle-" + ("dense" if mode == 1 else "block"))
    assert n >= min(3, len(golden["scenarios"][sc["name"]]["trace"]))


def test_oracle_thread_count_does_not_change_the_fit(c1):
    """The reference requires its KATs to hold for any thread count
    (c++/tests/lm_solver.cpp:27, CI OMP_NUM_THREADS=8)."""
    sc = scenarios.C1_SCENARIOS[0]
    a, _ = scenarios.run(sc, backends.oracle_solver(1), c1, n_threads=1)
    b, _ = scenarios.run(sc, backends.oracle_solver(1), c1, n_threads=4)
    np.testing.assert_allclose(a["pars"], b["pars"], rtol=1e-13)


def test_oracle_access_functions(c1, golden):
    """c++/tests/lm_solver.cpp:222-242 (checksums @1e-13, dof = 195)."""
    sc = scenarios.C1_SCENARIOS[0]
    s, lam, _, _ = scenarios.build(sc, backends.oracle_solver(1), c1)
    np.testing.assert_allclose(s.getValue(c1["fix_d"][2]), 2.960644529912441, rtol=1e-13)
    s.fit(lam)
    assert s.degreesOfFreedom() == 195
    np.testing.assert_allclose(s.getValue(c1["fix_d"][2]), 36.39905496310919, rtol=1e-13)
    np.testing.assert_allclose(np.sum(s.getJacobian()), 353.6485673748526, rtol=1e-13)
    np.testing.assert_allclose(np.sum(s.getJTJ()[:5]), 580.3488115472484, rtol=1e-13)
    np.testing.assert_allclose(np.sum(s.getDTD()[:5]), 34340.67196549198, rtol=1e-13)
    np.testing.assert_allclose(np.sum(s.getLeftSide()[:5]), 614.6894835127404, rtol=1e-13)
    np.testing.assert_allclose(np.sum(s.getRightSide()[:5]), 4410.585412402701, rtol=1e-13)
    np.testing.assert_allclose(np.sum(s.getResiduals()), 213.3530475167945, rtol=1e-13)
    # c++/tests/lm_solver.cpp:244-276: JTJ^-1 JTJ = 1
    n = 5
    prod = s.getInvJTJ().reshape(n, n) @ s.getJTJ().reshape(n, n)
    np.testing.assert_allclose(prod, np.eye(n), atol=1e-12)


@pytest.mark.skipif(not backends.has_ref(), reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("sc", scenarios.C1_SCENARIOS, ids=lambda s: s["name"])
def test_compiled_reference_reproduces_its_own_known_answers(sc, c1):
    """Pins the harness + spdlog stub: the unmodified reference behind the
    C-ABI gives the digits its own tests assert."""
    res, _ = scenarios.run(sc, backends.ref_solver(), c1)
    _check_kat(sc, res, c1)


@pytest.mark.skipif(not backends.has_ref(), reason="oracle/_ref not built (no /root/reference)")
def test_golden_vectors_are_what_the_reference_computes(c1, golden):
    for sc in scenarios.ALL_SCENARIOS:
        res, _ = scenarios.run(sc, backends.ref_solver(), c1)
        want = golden["scenarios"][sc["name"]]
        assert res["chi2"] == want["chi2"], sc["name"]
        assert res["pars"] == want["pars"], sc["name"]
        assert [t["decision"] for t in res["trace"]] == [t["decision"] for t in want["trace"]]


def test_oracle_error_codes_match_reference_exceptions(c1):
    """c++/tests/lm_solver.cpp:278-365"""
    import gadfit_b200 as g
    _run_exception_sections(backends.oracle_solver(), c1, g)


def _run_exception_sections(make, c1, g):
    d = c1["fix_d"]
    x1, y1, x2, y2 = c1["x_data_1"], c1["y_data_1"], c1["x_data_2"], c1["y_data_2"]

    def fresh():
        s = make("exponential")
        s.settings.iteration_limit = 4
        return s

    s = fresh()  # Incorrect number of calls
    s.addDataset(x1, y1)
    s.setPar(0, d[0], True, 0)
    with pytest.raises(g.LateAddDatasetCall):
        s.addDataset(x2, y2)
    s = fresh()  # Invalid data set index
    s.addDataset(x1, y1)
    s.addDataset(x2, y2)
    s.setPar(0, d[0], True, 0)
    with pytest.raises(g.SetParInvalidIndex):
        s.setPar(1, d[3], True, 2)
    s = fresh()  # Uninitialized fitting parameter
    s.addDataset(x1, y1)
    s.addDataset(x2, y2)
    s.setPar(0, d[0], True, 0)
    s.setPar(2, d[1], True, 0)
    s.setPar(0, d[4], True, 1)
    s.setPar(1, d[3], True)
    with pytest.raises(g.UninitializedParameter):
        s.fit(1.0)
    s = fresh()  # Too few data points
    s.addDataset(x1[:2].copy(), y1[:2].copy())
    s.addDataset(x2[:2].copy(), y2[:2].copy())
    s.setPar(0, d[0], True, 0)
    s.setPar(2, d[1], True, 0)
    s.setPar(0, d[4], True, 1)
    s.setPar(2, d[5], True, 1)
    s.setPar(1, d[3], True)
    with pytest.raises(g.NegativeDegreesOfFreedom):
        s.fit(1.0)
    s = fresh()  # No active parameters
    s.addDataset(x1, y1)
    s.addDataset(x2, y2)
    s.setPar(0, d[0], False, 0)
    s.setPar(2, d[1], False, 0)
    s.setPar(0, d[4], False, 1)
    s.setPar(2, d[5], False, 1)
    s.setPar(1, d[3], False)
    with pytest.raises(g.NoFittingParameters):
        s.fit(1.0)
    s = fresh()  # Active: bgr-0, I0-1 -> no global parameter
    s.addDataset(x1, y1)
    s.addDataset(x2, y2)
    s.setPar(0, d[0], False, 0)
    s.setPar(2, d[1], True, 0)
    s.setPar(0, d[4], True, 1)
    s.setPar(2, d[5], False, 1)
    s.setPar(1, d[12], False)
    with pytest.raises(g.NoGlobalParameters):
        s.fit(1.0)
