#!/bin/bash
# round-2 GPU pass 7: host-step completion word A/B, host-step tests, bench
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tasks.py -m gpu -q -k "step_host or rl_device" > gpurun_out/r02g_pytest.log 2>&1; tail -3 gpurun_out/r02g_pytest.log
OZL_HOST_FLAG=1 python profiles/e2e_breakdown.py > gpurun_out/r02g_e2e_flag1.json 2>&1; cat gpurun_out/r02g_e2e_flag1.json
OZL_HOST_FLAG=0 python profiles/e2e_breakdown.py > gpurun_out/r02g_e2e_flag0.json 2>&1; cat gpurun_out/r02g_e2e_flag0.json
python bench.py --steps 1000 --warmup 100 --no-side-configs --no-cpu-baseline > gpurun_out/r02g_bench_k1000.json 2> gpurun_out/r02g_bench_k1000.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/r02g_bench_k1000.json').read().strip().splitlines()[-1]); e=d['e2e']
print('e2e', e['value'], e['value_with_explicit_copies'], e['value_pipelined_two_halves'], e['per_rank'])"
