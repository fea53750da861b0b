"""Randomised parity: seeded random global-fit set-ups on synthetic
exponential / Gaussian data -- number of data sets, ragged sizes, which
parameters are shared / local / fixed (per data set), errors or none, lambda,
damping options, loss function, geodesic acceleration -- run through the
compiled reference, the oracle restatement and the device.

CPU: oracle vs unmodified reference.  GPU: device vs oracle.  Same checks as
the golden-vector tests (scenarios.compare_to_golden): identical accept/reject
sequence and lambda schedule, chi2 per trial, parameters and chi2 to 1e-10.
"""
import numpy as np
import pytest

import backends
import scenarios

N_CASES = 24
MODELS = {"exponential": 3, "gauss_linear_bg": 6}


def make_case(seed: int) -> dict:
    rng = np.random.default_rng(1000 + seed)
    model = "exponential" if seed % 3 else "gauss_linear_bg"
    n_pars = MODELS[model]
    n_sets = int(rng.integers(1, 5))
    sizes = [int(rng.integers(n_pars + 3, 120)) for _ in range(n_sets)]
    if model == "exponential":
        truth = lambda: np.array([rng.uniform(20, 150), rng.uniform(8, 30), rng.uniform(2, 20)])  # noqa: E731
        f = lambda p, x: p[0] * np.exp(-x / p[1]) + p[2]  # noqa: E731
        xs = [np.sort(rng.uniform(0.5, 60.0, n)) for n in sizes]
    else:
        truth = lambda: np.array([rng.uniform(1, 5), rng.uniform(-1, 1), rng.uniform(0.5, 1.2),  # noqa: E731
                                  rng.uniform(-0.2, 0.2), rng.uniform(0, 1), rng.uniform(-0.1, 0.1)])
        f = lambda p, x: (p[0] * np.exp(-((x - p[1]) / p[2]) ** 2) * (1 + p[3] * (x - p[1]) / p[2])  # noqa: E731
                          + p[4] + p[5] * x)
        xs = [np.sort(rng.uniform(-4.0, 4.0, n)) for n in sizes]
    # roles: each parameter is shared, local or fixed; at least one shared
    # when there are several data sets, at least one active overall
    roles = rng.choice(["shared", "local", "fixed"], size=n_pars, p=[0.4, 0.4, 0.2])
    if n_sets > 1 and "shared" not in roles:
        roles[int(rng.integers(n_pars))] = "shared"
    if all(r == "fixed" for r in roles):
        roles[0] = "shared"
    base = truth()
    pars, ys, errs = [], [], []
    with_errors = bool(rng.integers(2))
    for s in range(n_sets):
        p = truth()
        for k, r in enumerate(roles):
            if r != "local":
                p[k] = base[k]
        sigma = rng.uniform(0.5, 2.0, sizes[s]) * (0.02 * np.max(np.abs(f(p, xs[s]))) + 1e-3)
        ys.append(f(p, xs[s]) + sigma * rng.standard_normal(sizes[s]))
        errs.append(sigma if with_errors else None)
        pars.append(p)
    start = [p * rng.uniform(0.85, 1.15, n_pars) + 0.01 for p in pars]
    # a local parameter may additionally be fixed in individual data sets
    local_fixed = {(s, k) for s in range(n_sets) for k, r in enumerate(roles)
                   if r == "local" and n_sets > 1 and rng.random() < 0.15}
    settings = dict(iteration_limit=int(rng.integers(2, 9)),
                    lambda_incs=int(rng.integers(1, 4)),
                    lambda_up=float(rng.choice([10.0, 5.0, 3.0])),
                    lambda_down=float(rng.choice([10.0, 5.0, 2.0])),
                    damp_max=bool(rng.integers(2)),
                    loss=int(rng.choice([0, 0, 1, 2])),
                    acceleration_threshold=float(rng.choice([-1.0, 0.75, 0.9])))
    return dict(seed=seed, model=model, n_pars=n_pars, xs=xs, ys=ys, errs=errs, roles=list(roles),
                start=start, local_fixed=local_fixed, settings=settings,
                lambda0=float(rng.choice([10.0, 1.0, 0.1, 1e-3])))


def run_case(case: dict, make) -> dict:
    s = make(case["model"])
    for x, y, e in zip(case["xs"], case["ys"], case["errs"]):
        if e is None:
            s.addDataset(x, y)
        else:
            s.addDataset(x, y, e)
    s._keep = case
    n_sets = len(case["xs"])
    for k, role in enumerate(case["roles"]):
        if role == "shared":
            s.setPar(k, float(case["start"][0][k]), True)
        elif role == "fixed":
            s.setPar(k, float(case["start"][0][k]), False)
        else:
            for d in range(n_sets):
                s.setPar(k, float(case["start"][d][k]), (d, k) not in case["local_fixed"], d)
    for key, val in case["settings"].items():
        setattr(s.settings, key, val)
    s.fit(case["lambda0"])
    return dict(name=f"random_{case['seed']}", chi2=s.chi2(),
                pars=[[s.getParValue(k, d) for k in range(case["n_pars"])] for d in range(n_sets)],
                trace=s.trace(), iterations=s.iterations(), dof=s.degreesOfFreedom())


def usable(rec: dict) -> bool:
    """A random start may send the reference itself to NaN (rejected steps
    only); such a case still has to agree on the decisions, which the
    comparison checks, but there are no meaningful numbers to compare."""
    return np.isfinite(rec["chi2"])


@pytest.mark.skipif(not backends.has_ref(), reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("seed", range(N_CASES))
def test_oracle_matches_compiled_reference_on_random_setups(seed):
    case = make_case(seed)
    want = run_case(case, backends.ref_solver())
    got = run_case(case, backends.oracle_solver(1))
    assert usable(want) == usable(got)
    if usable(want):
        scenarios.compare_to_golden(got, want, rtol=1e-10, label="oracle-dense")
    got_block = run_case(case, backends.oracle_solver(2))  # block-arrow solve
    if usable(want):
        scenarios.compare_to_golden(got_block, want, rtol=1e-10, label="oracle-block")


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(N_CASES))
def test_device_matches_oracle_on_random_setups(seed):
    case = make_case(seed)
    want = run_case(case, backends.oracle_solver(1))
    got = run_case(case, backends.product_solver())
    assert usable(want) == usable(got)
    if usable(want):
        scenarios.compare_to_golden(got, want, rtol=1e-10, label="device-vs-oracle")
