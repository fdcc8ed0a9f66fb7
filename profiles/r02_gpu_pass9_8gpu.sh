#!/bin/bash
# round-2 GPU pass 9 (N GPUs): topology + the driver's scaling command at K=20 after the device-side rendezvous / NUMA memory policy
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
bash profiles/topology.sh > gpurun_out/r02i_topology.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --no-side-configs > gpurun_out/r02i_bench_${N}gpu_k20.json 2> gpurun_out/r02i_bench_${N}gpu_k20.err; echo rc=$?
python bench.py --steps 20 --warmup 5 --no-side-configs --no-cpu-baseline > gpurun_out/r02i_bench_1gpu_k20.json 2> gpurun_out/r02i_bench_1gpu_k20.err; echo rc=$?
