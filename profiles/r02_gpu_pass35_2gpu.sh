#!/bin/bash
# round-2 GPU pass 36 (N GPUs): what does a multi-process run cost WITHOUT any collective in the timed graph?
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
for C in peer-side peer peer-side; do timeout 600 $T bench.py --gpus $N --steps 20 --warmup 5 --no-side-configs --no-e2e --no-cpu-baseline --metrics-collective $C >> gpurun_out/r02ai_bench_${N}gpu_k20_$C.json 2> gpurun_out/r02ai_bench.err; done
for C in peer-side; do timeout 600 $T bench.py --gpus $N --steps 1000 --warmup 100 --no-side-configs --no-e2e --no-cpu-baseline --metrics-collective $C >> gpurun_out/r02ai_bench_${N}gpu_k1000_$C.json 2>> gpurun_out/r02ai_bench.err; done
