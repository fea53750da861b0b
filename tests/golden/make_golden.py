"""Generate tests/golden/golden_ref.json from the UNMODIFIED reference.

Run in the build container (needs oracle/_ref/libgadfit_ref.so, i.e.
/root/reference mounted and `make -C oracle ref` done):

    python tests/golden/make_golden.py

Every scenario of tests/scenarios.py is run through the compiled reference
(single thread, plain -O2 build, in-tree linear algebra) and the resulting
chi2, parameters, iteration count, degrees of freedom and the complete
accept/reject trace (lambda, new chi2, old chi2, decision per trial step --
the reference's own debug output, lm_solver.cpp:457-460) are stored.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import backends  # noqa: E402
import scenarios  # noqa: E402


def main():
    fx = json.load(open(os.path.join(HERE, "c1_fixture.json")))
    c1 = {k: (np.array(v) if isinstance(v, list) else v) for k, v in fx.items()}
    out = {"_generator": "tests/golden/make_golden.py",
           "_reference": "raullaasner/gadfit v2.0.1 c++ sources, g++ -O2 -fopenmp, lapack_fallback.cpp, 1 thread",
           "scenarios": {}}
    for sc in scenarios.ALL_SCENARIOS:
        res, solver = scenarios.run(sc, backends.ref_solver(), c1, n_threads=1)
        if sc["kind"] == "c1" and sc["name"] == "indexing_all":
            # "Access functions" test, c++/tests/lm_solver.cpp:222-242
            res["sums"] = dict(
                jacobian=float(np.sum(solver.getJacobian())),
                jtj5=float(np.sum(solver.getJTJ()[:5])),
                dtd5=float(np.sum(solver.getDTD()[:5])),
                left5=float(np.sum(solver.getLeftSide()[:5])),
                right5=float(np.sum(solver.getRightSide()[:5])),
                residuals=float(np.sum(solver.getResiduals())),
                jtj=solver.getJTJ().tolist(), dtd=solver.getDTD().tolist(),
                right=solver.getRightSide().tolist(), left=solver.getLeftSide().tolist(),
                inv_jtj=solver.getInvJTJ().tolist())
        if sc["kind"] == "workload":
            w = solver._scenario_keep[0]
            res["input_sha256"] = hashlib.sha256(
                np.concatenate([w.x, w.y, w.sigma, w.start.ravel()]).tobytes()).hexdigest()
        out["scenarios"][sc["name"]] = res
        print(sc["name"], res["chi2"], res["iterations"], len(res["trace"]))
    json.dump(out, open(os.path.join(HERE, "golden_ref.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
