#!/bin/bash
# round-2 GPU pass 38 (2 GPUs): bench falls back to the NCCL all-reduce when the ranks cannot map each other's memory; N = 2 record
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
OZL_XCHG_FAIL_IPC=1 timeout 600 $T bench.py --gpus $N --steps 20 --warmup 5 --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r02ak_bench_${N}gpu_fallback.json 2> gpurun_out/r02ak_fallback.err; echo rc=$?
grep "\[bench\]" gpurun_out/r02ak_fallback.err | head -3
timeout 600 $T bench.py --gpus $N --steps 20 --warmup 5 --no-side-configs > gpurun_out/r02ak_bench_${N}gpu_k20.json 2> gpurun_out/r02ak_bench.err; echo rc=$?
