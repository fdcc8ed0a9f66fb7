#!/bin/bash
# round-2 GPU pass 34 (N GPUs): metrics exchange with every load issued up front -- tests, check vs NCCL, bench at the driver's flags
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r02ag_pytest.log 2>&1; tail -5 gpurun_out/r02ag_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 300 $T benchmarks/peer_metrics_check.py > gpurun_out/r02ag_peer_check_${N}gpu.json 2> gpurun_out/r02ag_peer_check.err; cat gpurun_out/r02ag_peer_check_${N}gpu.json
for i in 1 2 3; do timeout 600 $T bench.py --gpus $N --steps 20 --warmup 5 --no-side-configs --no-e2e --no-cpu-baseline >> gpurun_out/r02ag_bench_${N}gpu_k20.json 2> gpurun_out/r02ag_bench.err; done
timeout 600 $T bench.py --gpus $N --steps 1000 --warmup 100 --no-side-configs --no-e2e --no-cpu-baseline >> gpurun_out/r02ag_bench_${N}gpu_k1000.json 2>> gpurun_out/r02ag_bench.err
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --no-side-configs --no-cpu-baseline --no-e2e >> gpurun_out/r02ag_bench_1gpu_k20.json 2>> gpurun_out/r02ag_bench.err; done
timeout 300 python bench.py --steps 1000 --warmup 100 --no-side-configs --no-cpu-baseline --no-e2e >> gpurun_out/r02ag_bench_1gpu_k1000.json 2>> gpurun_out/r02ag_bench.err
