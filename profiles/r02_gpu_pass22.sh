#!/bin/bash
# round-2 GPU pass 22: covariance-tile bulk copies issued by one lane per warp
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_config3.py tests/test_gpu_tasks.py -m gpu -q -x -k "ekf or config3 or chained or one_launch" > gpurun_out/r02u_pytest.log 2>&1; tail -4 gpurun_out/r02u_pytest.log
O=gpurun_out/r02u_config3.jsonl; : > $O
for i in 1 2 3; do timeout 300 python profiles/time_config3.py >> $O 2>gpurun_out/r02u.err; done
OZL_EKF_BLOCK=128 timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02u.err
timeout 300 python profiles/time_config3.py 262144 100 >> $O 2>>gpurun_out/r02u.err
timeout 300 python profiles/time_config3.py 4099 200 >> $O 2>>gpurun_out/r02u.err
cat $O
