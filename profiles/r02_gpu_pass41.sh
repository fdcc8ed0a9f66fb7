#!/bin/bash
# round-2 GPU pass 41: e2e of the driver's command with 5 vs 64 untimed warm-up steps of the host path
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for W in 5 64 5 64 256; do OZL_BENCH_E2E_WARMUP=$W python bench.py --steps 20 --warmup 5 --no-side-configs --no-cpu-baseline >> gpurun_out/r02an_e2e_warmup_$W.json 2>> gpurun_out/r02an.err; done
