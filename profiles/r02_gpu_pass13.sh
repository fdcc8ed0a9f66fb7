#!/bin/bash
# round-2 GPU pass 13: chaining restricted to multi-wave grids; CTA size 64 / 96 / 128; PV loop unroll 1 / 3 / 9
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_config3.py -m gpu -q -x > gpurun_out/r02m_pytest.log 2>&1; tail -5 gpurun_out/r02m_pytest.log
O=gpurun_out/r02m_config3.jsonl; : > $O
timeout 300 python profiles/time_config3.py >> $O 2>gpurun_out/r02m_config3.err
for B in 64 96; do OZL_EKF_BLOCK=$B timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02m_config3.err; done
for V in pvu9 pvu1; do OUZELUM_B200_LIB=$PWD/scratch/variants/lib_$V.so timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02m_config3.err; done
timeout 300 python profiles/time_config3.py 262144 100 >> $O 2>>gpurun_out/r02m_config3.err
OZL_EKF_BLOCK=64 timeout 300 python profiles/time_config3.py 262144 100 >> $O 2>>gpurun_out/r02m_config3.err
OZL_EKF_BLOCK=64 OZL_EKF_CHAIN=2 timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02m_config3.err
cat $O; tail -5 gpurun_out/r02m_config3.err
