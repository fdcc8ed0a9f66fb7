#!/bin/bash
# round-2 GPU pass 30: warp-private shared-memory tiles in the fused kernel (per-warp tensor-map copies, no block barriers after the prologue)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_config3.py tests/test_gpu_tasks.py -m gpu -q > gpurun_out/r02ac_pytest.log 2>&1; tail -4 gpurun_out/r02ac_pytest.log
O=gpurun_out/r02ac_config3.jsonl; : > $O
for i in 1 2; do
timeout 300 python profiles/time_config3.py >> $O 2>gpurun_out/r02ac.err
OZL_EKF_BLOCK=128 timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02ac.err
OZL_EKF_BLOCK=64 timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02ac.err
done
OZL_EKF_BLOCK=256 timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02ac.err
OZL_EKF_TMAP=0 timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02ac.err
timeout 300 python profiles/time_config3.py 262144 100 >> $O 2>>gpurun_out/r02ac.err
cat $O
