#!/bin/bash
# round-2 GPU pass 4: re-run the three fixed tests, full ncu capture of the fused EKFLeeLanded kernel, register variants
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_companions.py tests/test_gpu_dr.py tests/test_gpu_tasks.py -m gpu -q -x -k "arbiter or uniform_schedules or graphed_rollout_equals" > gpurun_out/r02d_pytest.log 2>&1; tail -3 gpurun_out/r02d_pytest.log
python profiles/time_config3.py > gpurun_out/r02d_config3.jsonl 2>gpurun_out/r02d_config3.err
cat gpurun_out/r02d_config3.jsonl
ncu --set full --clock-control none --import-source on -k regex:ekf_lee_fused -s 70 -c 1 -o gpurun_out/r02d_ekf_fused python profiles/time_config3.py 65536 2 > gpurun_out/r02d_ncu.log 2>&1; tail -2 gpurun_out/r02d_ncu.log
