#!/bin/bash
# round-2 GPU pass 39 (4 GPUs): the driver's scaling command at N = 4 with the final binary, both arms
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-4}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $T bench.py --gpus $N --steps 20 --warmup 5 --no-side-configs > gpurun_out/r02al_bench_${N}gpu_k20.json 2> gpurun_out/r02al_bench.err; echo rc=$?
timeout 600 $T bench.py --gpus $N --steps 20 --warmup 5 --impl reference > gpurun_out/r02al_bench_${N}gpu_ref.json 2>/dev/null; echo rc=$?
