#!/bin/bash
# round-2 GPU pass 46: step kernel with the episode-statistics reductions issued before the stores
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r02az_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02az_pytest.log; tail -3 gpurun_out/r02az_pytest.log
python profiles/time_sizes.py 16384 65536 131072 > gpurun_out/r02az_sizes.jsonl 2> gpurun_out/r02az.err
for i in 1 2 3; do python bench.py --steps 20 --warmup 5 --no-side-configs --no-cpu-baseline --no-e2e >> gpurun_out/r02az_bench_k20.json 2>> gpurun_out/r02az.err; done
python bench.py --steps 1000 --warmup 100 --no-side-configs --no-cpu-baseline --no-e2e >> gpurun_out/r02az_bench_k1000.json 2>> gpurun_out/r02az.err
