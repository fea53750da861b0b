"""Developer probe (GPU): first point at which the fused small-problem path
and the stage-wise path differ on a C1 scenario.  python tests/dev/fused_vs_staged.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import json

import backends
import scenarios

fx = json.load(open(os.path.join(os.path.dirname(HERE), "golden", "c1_fixture.json")))
c1 = {k: (np.array(v) if isinstance(v, list) else v) for k, v in fx.items()}
name = sys.argv[1] if len(sys.argv) > 1 else "indexing_all"
sc = [s for s in scenarios.ALL_SCENARIOS if s["name"] == name][0]
for limit in (1, 2, 3):
    out = {}
    for fused in ("1", "0"):
        os.environ["GADFIT_B200_FUSED"] = fused
        s, lam, n_sets, n_pars = scenarios.build(sc, backends.product_solver(), c1)
        s.settings.iteration_limit = limit
        s.fit(lam)
        out[fused] = dict(
            pars=np.array([[s.getParValue(k, i) for k in range(n_pars)] for i in range(n_sets)]),
            chi2=np.array([s.chi2()]), jtj=s.getJTJ(), dtd=s.getDTD(), rhs=s.getRightSide(),
            lhs=s.getLeftSide(), trace=np.array([[t["new_chi2"], t["old_chi2"], t["lambda_"]] for t in s.trace()]))
    for key in out["1"]:
        a, b = out["1"][key], out["0"][key]
        same = a.shape == b.shape and np.array_equal(a, b)
        print(limit, key, "identical" if same else f"DIFFERENT max rel {np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)):.3e}")
