#!/bin/bash
# round-2 GPU pass 6: full parity suite, config 3, e2e breakdown, bench at the driver's K
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; tail -6 gpurun_out/r02f_pytest.log
python profiles/time_config3.py > gpurun_out/r02f_config3.jsonl 2>gpurun_out/r02f_config3.err; cat gpurun_out/r02f_config3.jsonl
python profiles/e2e_breakdown.py > gpurun_out/r02f_e2e_breakdown.json 2>&1; cat gpurun_out/r02f_e2e_breakdown.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_k20.json 2> gpurun_out/r02f_bench_k20.err; echo rc=$?
