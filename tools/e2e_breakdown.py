"""Developer script (run under gpurun): where the end-to-end time of bench.py's
C2 arm goes -- host-side description, H2D upload, LM iterations -- next to the
raw pinned H2D bandwidth of the box."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gadfit_b200 import LMsolver
from gadfit_b200 import workloads as W

torch.cuda.set_device(0)
pinned = []
def alloc(n):
    t = torch.empty(n, dtype=torch.float64, pin_memory=True); pinned.append(t); return t.numpy()
w = W.WORKLOADS["C2"](n_sets=64, n_points=1_000_000, alloc=alloc, owned=range(64), iteration_limit=10)
# raw H2D
big = torch.empty(64_000_000, dtype=torch.float64, pin_memory=True)
dev = torch.empty(64_000_000, dtype=torch.float64, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dev.copy_(big, non_blocking=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("raw pinned H2D 512 MB: %.2f ms = %.1f GB/s" % (dt * 1e3, 0.512 / dt))
for rep in range(4):
    s = LMsolver(w.model); s.setDevice(0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    w.setup(s, pointers=False); s.settings.iteration_limit = 10
    t1 = time.perf_counter()
    s.upload(); torch.cuda.synchronize()
    t2 = time.perf_counter()
    s.fit(w.lambda0); torch.cuda.synchronize()
    t3 = time.perf_counter()
    out = s.chi2()
    t4 = time.perf_counter()
    del s; torch.cuda.synchronize()
    t5 = time.perf_counter()
    print("rep %d: describe %.2f ms, upload %.2f ms (%.1f GB/s), fit %.2f ms, chi2 %.2f ms, destroy %.2f ms, total %.2f" % (
        rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, 1.536 / (t2 - t1), (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t5 - t4) * 1e3, (t4 - t0) * 1e3))

# the same without the explicit upload(): fit() pipelines the copies with the first Jacobian sweep
for rep in range(4):
    s = LMsolver(w.model); s.setDevice(0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    w.setup(s, pointers=False); s.settings.iteration_limit = 10
    t1 = time.perf_counter()
    s.fit(w.lambda0); torch.cuda.synchronize()
    t3 = time.perf_counter()
    out = s.chi2()
    t4 = time.perf_counter()
    del s; torch.cuda.synchronize()
    print("rep %d: describe %.2f ms, fit incl. pipelined upload %.2f ms, chi2 %.2f ms, total %.2f (chi2 = %.10g)" % (
        rep, (t1 - t0) * 1e3, (t3 - t1) * 1e3, (t4 - t3) * 1e3, (t4 - t0) * 1e3, out))
