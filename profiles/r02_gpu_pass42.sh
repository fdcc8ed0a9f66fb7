#!/bin/bash
# round-2 GPU pass 42: e2e with the stepping thread pinned to one core vs free to migrate (K = 20, 3 runs each)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for P in 1 0 1 0 1 0; do OZL_BENCH_PIN_THREAD=$P python bench.py --steps 20 --warmup 5 --no-side-configs --no-cpu-baseline >> gpurun_out/r02ap_e2e_pin_$P.json 2>> gpurun_out/r02ap.err; done
