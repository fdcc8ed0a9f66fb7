#!/bin/bash
# round-2 GPU pass 1: parity tests, bench (both arms), launch list, full captures at 16k and 131k
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_k20.json 2> gpurun_out/r02a_bench_k20.err; echo rc=$?
python bench.py --steps 1000 --warmup 100 --no-side-configs --no-cpu-baseline > gpurun_out/r02a_bench_k1000.json 2> gpurun_out/r02a_bench_k1000.err; echo rc=$?
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02a_bench_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/r02a_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:quad_step_kernel -s 80 -c 1 -o gpurun_out/r02a_quad_step_16k python profiles/prof_step.py 16384 > gpurun_out/r02a_ncu16k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:quad_step_kernel -s 30 -c 1 -o gpurun_out/r02a_quad_step_131k python profiles/prof_step.py 131072 > gpurun_out/r02a_ncu131k.log 2>&1
python profiles/time_sizes.py > gpurun_out/r02a_sizes.json 2>&1
