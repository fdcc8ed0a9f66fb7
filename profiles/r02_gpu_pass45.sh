#!/bin/bash
# round-2 GPU pass 45: last check of the final tree + capture of the TMA step kernel at 262144 envs (its new default range)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r02ay_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ay_pytest.log; tail -3 gpurun_out/r02ay_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ay_smoke.log 2>&1; tail -1 gpurun_out/r02ay_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02ay_bench_k20.json 2> gpurun_out/r02ay_bench.err; echo rc=$?
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02ay_bench_ref.json 2>/dev/null; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:quad_step_tma -s 5 -c 1 -o gpurun_out/r02ay_quad_step_tma_262144 python profiles/prof_step.py 262144 20 > gpurun_out/r02ay_ncu.log 2>&1
