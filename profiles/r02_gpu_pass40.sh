#!/bin/bash
# round-2 GPU pass 40: last check of the final tree -- parity suite, smoke, bench at the driver's flags
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r02am_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02am_pytest.log; tail -4 gpurun_out/r02am_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02am_smoke.log 2>&1; tail -1 gpurun_out/r02am_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02am_bench_k20.json 2> gpurun_out/r02am_bench_k20.err; echo rc=$?
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02am_bench_ref.json 2>/dev/null; echo rc=$?
