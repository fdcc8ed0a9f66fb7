#!/bin/bash
# round-2 GPU pass 5: full parity suite, config-3 variants after the load hoists / EKF rewrite
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02e_pytest.log 2>&1; tail -6 gpurun_out/r02e_pytest.log
python profiles/time_config3.py > gpurun_out/r02e_config3.jsonl 2>gpurun_out/r02e_config3.err
for V in pvunroll3 philox_inline philox_inline_unroll3; do OUZELUM_B200_LIB=$PWD/scratch/variants/lib_$V.so python profiles/time_config3.py >> gpurun_out/r02e_config3.jsonl 2>>gpurun_out/r02e_config3.err; done
cat gpurun_out/r02e_config3.jsonl
