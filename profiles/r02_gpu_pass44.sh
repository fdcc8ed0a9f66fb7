#!/bin/bash
# round-2 GPU pass 44: TMA-pipelined step kernel from 166k envs (was 400k) -- parity suite, size sweep, bench
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r02au_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02au_pytest.log; tail -3 gpurun_out/r02au_pytest.log
python profiles/time_sizes.py 131072 163840 180224 196608 262144 393216 524288 1048576 > gpurun_out/r02au_sizes.jsonl 2> gpurun_out/r02au.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02au_bench_k20.json 2>> gpurun_out/r02au.err; echo rc=$?
