"""Full-size parity record: BASELINE config 2 (64 data sets x 1M points, 20
shared parameters) fitted by the CUDA library and by the UNMODIFIED reference
(oracle/_ref, all host threads) on identical inputs.  Writes
gpurun_out/full_size_parity_c2.json (copied to profiles/ once produced).

    gpurun -- python tests/dev/full_size_parity.py [n_sets] [n_points] [iterations]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gadfit_b200 import LMsolver, workloads as W  # noqa: E402
from gadfit_b200._capi import Api  # noqa: E402

n_sets = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_points = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
w = W.c2(n_sets=n_sets, n_points=n_points, iteration_limit=iters)

t0 = time.time()
g = LMsolver(w.model)
w.setup(g)
g.fit(w.lambda0)
t_gpu = time.time() - t0
gp = np.array([g.getParValue(k) for k in range(20)])
gtr = g.trace()
print("gpu done", t_gpu, flush=True)

ref_lib = os.path.join(ROOT, "oracle", "_ref", "libgadfit_ref_fast.so")
api = Api(ref_lib, "gadfit_ref_", product=False)
r = LMsolver(w.model, _api=api)
w.setup(r)
r.settings.n_threads = os.cpu_count()
t0 = time.time()
r.fit(w.lambda0)
t_ref = time.time() - t0
rp = np.array([r.getParValue(k) for k in range(20)])
rtr = r.trace()

out = dict(
    workload=f"C2: {n_sets} data sets x {n_points} points, 20 shared parameters, {iters} iterations",
    reference="unmodified reference c++ sources, g++ -O3 -march=x86-64-v3 -fopenmp, in-tree linear algebra, "
              f"{os.cpu_count()} threads",
    seconds=dict(gpu_fit_incl_upload=t_gpu, reference_fit=t_ref),
    reference_timings=r.timings(),
    iterations=dict(gpu=g.iterations(), reference=r.iterations()),
    decisions=dict(gpu=[t["decision"] for t in gtr], reference=[t["decision"] for t in rtr]),
    decisions_identical=[t["decision"] for t in gtr] == [t["decision"] for t in rtr],
    lambdas_identical=[t["lambda_"] for t in gtr] == [t["lambda_"] for t in rtr],
    chi2_trace_max_rel_diff=float(max(abs(a["new_chi2"] - b["new_chi2"]) / abs(b["new_chi2"])
                                      for a, b in zip(gtr, rtr))),
    final_chi2=dict(gpu=gtr[-1]["new_chi2"], reference=rtr[-1]["new_chi2"]),
    parameter_max_rel_diff=float(np.max(np.abs(gp - rp) / np.abs(rp))),
    parameters=dict(gpu=gp.tolist(), reference=rp.tolist()),
)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "full_size_parity_c2.json"), "w"), indent=1)
print(json.dumps({k: out[k] for k in ("seconds", "iterations", "decisions_identical", "lambdas_identical",
                                       "chi2_trace_max_rel_diff", "parameter_max_rel_diff")}))
