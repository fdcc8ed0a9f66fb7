#!/bin/bash
# round-2 GPU pass 24: covariance tile through ONE 2-D tensor-map copy per CTA (UTMALDG / UTMASTG) instead of 81 1-D bulk copies
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_config3.py tests/test_gpu_tasks.py -m gpu -q -k "ekf or config3 or chained or one_launch or checkpoint" > gpurun_out/r02w_pytest.log 2>&1; tail -4 gpurun_out/r02w_pytest.log
O=gpurun_out/r02w_config3.jsonl; : > $O
for i in 1 2; do
timeout 300 python profiles/time_config3.py >> $O 2>gpurun_out/r02w.err
OZL_EKF_TMAP=0 timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02w.err
done
OZL_EKF_BLOCK=128 timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02w.err
timeout 300 python profiles/time_config3.py 262144 100 >> $O 2>>gpurun_out/r02w.err
cat $O; tail -3 gpurun_out/r02w.err
