"""Developer script (under gpurun): the Fortran known answers of
tests/test_gadf_driver.py with each Jacobian kernel."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import backends  # noqa: E402
import test_gadf_driver as T  # noqa: E402

fx = json.load(open(os.path.join(ROOT, "tests", "golden", "fortran_fixture.json")))
fx = {k: (np.array(v) if isinstance(v, list) else v) for k, v in fx.items()}
which = sys.argv[1] if len(sys.argv) > 1 else "double"
make = backends.product_solver() if (len(sys.argv) < 3 or sys.argv[2] == "device") else backends.oracle_solver(1)
s = {"double": T.run_integral_double, "single": T.run_integral_single}[which](make, fx)
print(os.environ.get("GADFIT_B200_JAC"), which, "par0 %.16g par1 %.16g" % (s.getParValue(0), s.getParValue(1)),
      "iterations", s.iterations(), [(t["decision"], "%.15g" % t["new_chi2"]) for t in s.trace()])
