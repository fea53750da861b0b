"""FP64 pipe micro-benchmarks (run under gpurun); see csrc/microbench.cu."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gadfit_b200._capi import product_api
api = product_api()
names = {0: "DFMA, 8 chains x 64 warps/SM", 1: "DMMA m8n8k4, 8 chains x 64 warps/SM",
         2: "DMMA + DFMA interleaved", 3: "DMMA, 1 warp per sub-partition, 6 chains",
         4: "DMMA, 1 warp per sub-partition, 1 chain (latency)",
         5: "DFMA, 1 warp per sub-partition, 1 chain (latency)"}
for which, name in names.items():
    out = C.c_double()
    st = api.microbench(which, C.byref(out))
    print(f"{which} {name}: status {st}  {out.value / 1e12:.3f} TFLOP/s", flush=True)
