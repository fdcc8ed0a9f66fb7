#!/bin/bash
# round-2 GPU pass 16: warp-cooperative PV fixes + grouped-row correct / fused process noise -- full parity suite, config-3 timing A/B
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02p_pytest.log 2>&1; tail -15 gpurun_out/r02p_pytest.log
O=gpurun_out/r02p_config3.jsonl; : > $O
for i in 1 2; do
timeout 300 python profiles/time_config3.py >> $O 2>gpurun_out/r02p_config3.err
OUZELUM_B200_LIB=$PWD/scratch/variants/lib_coop0.so timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02p_config3.err
done
OZL_EKF_BLOCK=128 timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02p_config3.err
cat $O
