#!/bin/bash
# round-2 GPU pass 8 (N GPUs): the driver's scaling command at K=20, plus K=1000 for reference
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02h_bench_${N}gpu_k20.json 2> gpurun_out/r02h_bench_${N}gpu_k20.err; echo rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 1000 --warmup 100 --no-side-configs > gpurun_out/r02h_bench_${N}gpu_k1000.json 2> gpurun_out/r02h_bench_${N}gpu_k1000.err; echo rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r02h_bench_${N}gpu_ref.json 2>/dev/null; echo rc=$?
python bench.py --steps 20 --warmup 5 --no-side-configs --no-cpu-baseline > gpurun_out/r02h_bench_1gpu_k20.json 2> gpurun_out/r02h_bench_1gpu_k20.err; echo rc=$?
for f in gpurun_out/r02h_bench_${N}gpu_k20.json gpurun_out/r02h_bench_${N}gpu_k1000.json gpurun_out/r02h_bench_1gpu_k20.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); e=d['e2e']
print('$f', d['n_gpus'], 'value', d['value'], 'ms/step', d['ms_per_step'], 'per rank', d['ms_per_rank'], 'nccl', d['nccl_allreduce_in_timed_graph'])
print('  e2e', e['value'], [ (r['rank'], round(r['us_per_step_device'],1), round(r['host_p99_us'],1)) for r in e['per_rank']], e['numa_cpus_rank0'])"; done
