#!/bin/bash
# round-2 GPU pass 29: code-size variants of the fused kernel (cold fallback loops rolled; reset-path draws out of line; all draws out of line)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r02ab_config3.jsonl; : > $O
for i in 1 2; do
timeout 300 python profiles/time_config3.py >> $O 2>gpurun_out/r02ab.err
OUZELUM_B200_LIB=$PWD/scratch/variants/lib_colddraw.so timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02ab.err
OUZELUM_B200_LIB=$PWD/scratch/variants/lib_philox_ni.so timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02ab.err
done
cat $O
