#!/bin/bash
# round-2 GPU pass 12: tile-chained launches of the one-launch EKFLeeLanded step -- parity (graph replay == eager) and timing A/B
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_config3.py -m gpu -q -x > gpurun_out/r02l_pytest.log 2>&1; tail -15 gpurun_out/r02l_pytest.log
OZL_EKF_CHAIN=0 timeout 300 python profiles/time_config3.py > gpurun_out/r02l_config3.jsonl 2>gpurun_out/r02l_config3.err
OZL_EKF_CHAIN=1 timeout 300 python profiles/time_config3.py >> gpurun_out/r02l_config3.jsonl 2>>gpurun_out/r02l_config3.err
OZL_EKF_CHAIN=1 timeout 300 python profiles/time_config3.py 65536 1000 >> gpurun_out/r02l_config3.jsonl 2>>gpurun_out/r02l_config3.err
OZL_EKF_CHAIN=1 timeout 300 python profiles/time_config3.py 262144 100 >> gpurun_out/r02l_config3.jsonl 2>>gpurun_out/r02l_config3.err
OZL_EKF_CHAIN=0 timeout 300 python profiles/time_config3.py 262144 100 >> gpurun_out/r02l_config3.jsonl 2>>gpurun_out/r02l_config3.err
cat gpurun_out/r02l_config3.jsonl; tail -5 gpurun_out/r02l_config3.err
