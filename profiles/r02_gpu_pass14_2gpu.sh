#!/bin/bash
# round-2 GPU pass 14 (2 GPUs): NVLink peer-memory metrics exchange -- tests, torchrun check vs NCCL, bench A/B (peer vs nccl) at the
# driver's flags; plus the config-3 CTA-size A/B on GPU 0
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r02n_pytest.log 2>&1; tail -15 gpurun_out/r02n_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 benchmarks/peer_metrics_check.py > gpurun_out/r02n_peer_check_${N}gpu.json 2> gpurun_out/r02n_peer_check.err; cat gpurun_out/r02n_peer_check_${N}gpu.json; tail -5 gpurun_out/r02n_peer_check.err
for C in peer nccl peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --no-side-configs --no-e2e --metrics-collective $C >> gpurun_out/r02n_bench_${N}gpu_k20_$C.json 2> gpurun_out/r02n_bench_${N}gpu_$C.err; echo rc=$?
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-side-configs --no-cpu-baseline --no-e2e > gpurun_out/r02n_bench_1gpu_k20.json 2> gpurun_out/r02n_bench_1gpu.err; echo rc=$?
O=gpurun_out/r02n_config3.jsonl; : > $O
for B in 128 96 128 96; do OZL_EKF_BLOCK=$B timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02n_config3.err; done
OZL_EKF_BLOCK=96 OUZELUM_B200_LIB=$PWD/scratch/variants/lib_nreg136.so timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02n_config3.err
cat $O
