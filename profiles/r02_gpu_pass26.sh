#!/bin/bash
# round-2 GPU pass 26: cheaper bookkeeping in the cooperative fixes (no integer-division helper calls, no __fns loop); reset path out of line
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_config3.py tests/test_gpu_tasks.py tests/test_gpu_step.py -m gpu -q > gpurun_out/r02y_pytest.log 2>&1; tail -4 gpurun_out/r02y_pytest.log
O=gpurun_out/r02y_config3.jsonl; : > $O
for i in 1 2; do
timeout 300 python profiles/time_config3.py >> $O 2>gpurun_out/r02y.err
OUZELUM_B200_LIB=$PWD/scratch/variants/lib_coldreset.so timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02y.err
OUZELUM_B200_LIB=$PWD/scratch/variants/lib_coop0.so timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02y.err
done
cat $O
OUZELUM_B200_LIB=$PWD/scratch/variants/lib_coldreset.so timeout 600 python -m pytest tests/test_gpu_config3.py -m gpu -q > gpurun_out/r02y_pytest_cold.log 2>&1; tail -3 gpurun_out/r02y_pytest_cold.log
