#!/bin/bash
# round-2 GPU pass 47 (8 GPUs): the driver's scaling command with the final tree (K = 20, e2e included) + the one-GPU run on the same box
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $T bench.py --gpus $N --steps 20 --warmup 5 --no-side-configs --no-cpu-baseline > gpurun_out/r02ba_bench_${N}gpu_k20.json 2> gpurun_out/r02ba_bench.err; echo rc=$?
timeout 300 python bench.py --steps 20 --warmup 5 --no-side-configs --no-cpu-baseline > gpurun_out/r02ba_bench_1gpu_k20.json 2>> gpurun_out/r02ba_bench.err; echo rc=$?
