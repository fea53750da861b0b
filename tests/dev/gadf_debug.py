"""Developer script (run under gpurun): gadf driver, block-arrow case, device vs oracle per trial."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import backends
import gadfit_b200 as g
from gadfit_b200 import workloads as W
recs = []
for make in (backends.product_solver(), backends.oracle_solver(1)):
    w = W.WORKLOADS["C3"](n_sets=12, n_points=300)
    s = make(w.model)
    w.setup(s)
    s.gadf = g.GadfOptions(enabled=True, uphill=1, rel_error=1e-7, cos_phi=1e-9)
    s.settings.iteration_limit = 30
    s.settings.lambda_incs = 2
    s.settings.acceleration_threshold = 0.9
    s.fit(1.0)
    recs.append((s.trace(), s.stopReason(), s.iterations(), s.chi2()))
for tg, to in zip(recs[0][0], recs[1][0]):
    print("it %2d lam %.3e/%.3e new %.10e rel %.2e old rel %.2e dec %d %d" % (
        to["iteration"], tg["lambda_"], to["lambda_"], to["new_chi2"],
        abs(tg["new_chi2"] - to["new_chi2"]) / to["new_chi2"],
        abs(tg["old_chi2"] - to["old_chi2"]) / to["old_chi2"], tg["decision"], to["decision"]))
print(recs[0][1:], recs[1][1:])
