#!/bin/bash
# Round 2, GPU pass s: the kernel-variant agreement tests
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "variants_agree or staged_chi2 or phase_timers" > $OUT/r02s_pytest_variants.log 2>&1
tail -n 25 $OUT/r02s_pytest_variants.log
