#!/bin/bash
# round-2 GPU pass 10: full record of HEAD -- parity tests, smoke, bench (both arms, driver flags), launch list, sizes, config 3
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
tail -5 gpurun_out/r02j_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02j_smoke.log 2>&1; tail -1 gpurun_out/r02j_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02j_bench_k20.json 2> gpurun_out/r02j_bench_k20.err; echo rc=$?
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02j_bench_ref.json 2>/dev/null; echo rc=$?
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02j_launches.csv python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/r02j_ncu_launch.log 2>&1
python profiles/time_sizes.py > gpurun_out/r02j_sizes.json 2>&1
python profiles/time_config3.py > gpurun_out/r02j_config3.jsonl 2> gpurun_out/r02j_config3.err
python profiles/e2e_breakdown.py > gpurun_out/r02j_e2e_breakdown.json 2>&1
