#!/bin/bash
# round-2 GPU pass 20: ablation of the DATA MOVEMENT of the one-launch EKFLeeLanded step (timing only; 63 = all arithmetic removed)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r02s_config3_ablation.jsonl; : > $O
timeout 300 python profiles/time_config3.py >> $O 2>gpurun_out/r02s.err
for V in abl256 abl64 abl63 abl127 abl191 abl319 abl511; do OUZELUM_B200_LIB=$PWD/scratch/variants/lib_$V.so timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02s.err; done
cat $O
