#!/bin/bash
# round-2 GPU pass 15 (8 GPUs): the driver's scaling command with the NVLink peer-memory metrics exchange vs the NCCL all-reduce
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 benchmarks/peer_metrics_check.py > gpurun_out/r02o_peer_check_${N}gpu.json 2> gpurun_out/r02o_peer_check.err; cat gpurun_out/r02o_peer_check_${N}gpu.json
for C in peer nccl peer; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --no-side-configs --no-e2e --metrics-collective $C >> gpurun_out/r02o_bench_${N}gpu_k20_$C.json 2> gpurun_out/r02o_bench_${N}gpu_$C.err; echo rc=$?
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-side-configs --no-cpu-baseline --no-e2e > gpurun_out/r02o_bench_1gpu_k20.json 2> gpurun_out/r02o_bench_1gpu.err; echo rc=$?
