#!/bin/bash
# round-2 GPU pass 37 (8 GPUs): metrics exchange on a side stream forked inside the graph (bench default) vs in the chain
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
for C in peer peer-inline peer; do timeout 600 $T bench.py --gpus $N --steps 20 --warmup 5 --no-side-configs --no-e2e --no-cpu-baseline --metrics-collective $C >> gpurun_out/r02aj_bench_${N}gpu_k20_$C.json 2> gpurun_out/r02aj_bench.err; done
timeout 600 $T bench.py --gpus $N --steps 1000 --warmup 100 --no-side-configs --no-e2e --no-cpu-baseline >> gpurun_out/r02aj_bench_${N}gpu_k1000_peer.json 2>> gpurun_out/r02aj_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-side-configs --no-cpu-baseline --no-e2e >> gpurun_out/r02aj_bench_1gpu_k20.json 2>> gpurun_out/r02aj_bench.err
