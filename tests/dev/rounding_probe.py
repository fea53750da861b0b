"""Which side of the full-size C2 difference is rounding?  Fits the same C2
instance three ways -- CUDA library, unmodified reference (double, sequential
sums), and the oracle restatement with the shared block / J^T r / chi2
accumulated in long double ("extended") -- and reports the parameter distance
of the two double-precision solvers from the extended one.

    gpurun -- python tests/dev/rounding_probe.py [n_sets] [n_points] [iterations]
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gadfit_b200 import LMsolver, workloads as W  # noqa: E402
from gadfit_b200._capi import Api  # noqa: E402

n_sets = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n_points = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
w = W.c2(n_sets=n_sets, n_points=n_points, iteration_limit=iters)


def run(api=None, mode=None):
    s = LMsolver(w.model, _api=api)
    if mode is not None:
        api.lib.gadfit_oracle_set_mode.argtypes = [C.c_void_p, C.c_int]
        api.lib.gadfit_oracle_set_mode(s._h, mode)
    w.setup(s)
    s.settings.n_threads = os.cpu_count()
    s.fit(w.lambda0)
    return np.array([s.getParValue(k) for k in range(20)]), s.trace()


gpu, gtr = run()
ref, rtr = run(Api(os.path.join(ROOT, "oracle", "_ref", "libgadfit_ref_fast.so"), "gadfit_ref_", product=False))
ext, etr = run(Api(os.path.join(ROOT, "oracle", "_build", "libgadfit_oracle.so"), "gadfit_oracle_", product=False),
               mode=1 | 4)
rel = lambda a, b: np.abs(a - b) / np.abs(b)  # noqa: E731
out = dict(
    workload=f"C2: {n_sets} x {n_points} points, {iters} iterations",
    decisions_identical=[t["decision"] for t in gtr] == [t["decision"] for t in rtr] == [t["decision"] for t in etr],
    gpu_vs_extended=dict(max=float(rel(gpu, ext).max()), per_parameter=rel(gpu, ext).tolist()),
    reference_vs_extended=dict(max=float(rel(ref, ext).max()), per_parameter=rel(ref, ext).tolist()),
    gpu_vs_reference=dict(max=float(rel(gpu, ref).max())),
    final_chi2=dict(gpu=gtr[-1]["new_chi2"], reference=rtr[-1]["new_chi2"], extended=etr[-1]["new_chi2"]),
)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "rounding_probe_c2.json"), "w"), indent=1)
print(json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk != "per_parameter"})
                  for k, v in out.items()}))
