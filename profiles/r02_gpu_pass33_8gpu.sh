#!/bin/bash
# round-2 GPU pass 33 (N GPUs): the driver's scaling commands with the final binary -- K = 20 (e2e included), K = 1000, config 4
# (131072 envs per GPU = 1 Mi envs on 8 GPUs, metrics exchange every 16 steps), both arms, plus the one-GPU runs on the same box
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $T bench.py --gpus $N --steps 20 --warmup 5 --no-side-configs > gpurun_out/r02af_bench_${N}gpu_k20.json 2> gpurun_out/r02af_bench_${N}gpu_k20.err; echo rc=$?
timeout 600 $T bench.py --gpus $N --steps 20 --warmup 5 --impl reference > gpurun_out/r02af_bench_${N}gpu_ref.json 2> /dev/null; echo rc=$?
timeout 600 $T bench.py --gpus $N --steps 1000 --warmup 100 --no-side-configs --no-cpu-baseline --no-e2e > gpurun_out/r02af_bench_${N}gpu_k1000.json 2> gpurun_out/r02af_bench_${N}gpu_k1000.err; echo rc=$?
timeout 600 $T bench.py --gpus $N --steps 1000 --warmup 100 --envs 131072 --no-side-configs --no-cpu-baseline --no-e2e > gpurun_out/r02af_config4_${N}gpu_131072_per_gpu.json 2> gpurun_out/r02af_config4.err; echo rc=$?
timeout 300 python bench.py --steps 20 --warmup 5 --no-side-configs --no-cpu-baseline > gpurun_out/r02af_bench_1gpu_k20.json 2> gpurun_out/r02af_bench_1gpu.err; echo rc=$?
timeout 300 python bench.py --steps 1000 --warmup 100 --envs 131072 --no-side-configs --no-cpu-baseline --no-e2e > gpurun_out/r02af_config4_1gpu_131072.json 2> gpurun_out/r02af_config4_1gpu.err; echo rc=$?
