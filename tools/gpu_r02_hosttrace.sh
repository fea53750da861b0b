#!/bin/bash
# Developer pass (gpurun --gpus 2): host-side phases of fit() on config 5 with 2 ranks
set -u
OUT=gpurun_out
GADFIT_B200_TRACE_HOST=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 2 --workload C5 --steps 6 --warmup 3 --repeats 2 --no-cpu-baseline --no-e2e > $OUT/r02_hosttrace_C5_n2.json 2> $OUT/r02_hosttrace_C5_n2.err
grep "gadfit_b200 host" $OUT/r02_hosttrace_C5_n2.err | tail -24
