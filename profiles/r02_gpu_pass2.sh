#!/bin/bash
# round-2 GPU pass 2 (run with --gpus 2): parity tests, bench at the driver's K=20 and at K=1000 (1 and 2 GPUs), reference arm
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -3 gpurun_out/r02b_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_k20.json 2> gpurun_out/r02b_bench_k20.err; echo rc=$?
python bench.py --steps 1000 --warmup 100 --no-side-configs --no-cpu-baseline > gpurun_out/r02b_bench_k1000.json 2> gpurun_out/r02b_bench_k1000.err; echo rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02b_bench_2gpu_k20.json 2> gpurun_out/r02b_bench_2gpu_k20.err; echo rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1000 --warmup 100 --no-side-configs > gpurun_out/r02b_bench_2gpu_k1000.json 2> gpurun_out/r02b_bench_2gpu_k1000.err; echo rc=$?
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02b_bench_ref.json 2>/dev/null
