#!/bin/bash
# round-2 GPU pass 11: stage ablation of the one-launch EKFLeeLanded step (timing only) + fresh full capture
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python profiles/time_config3.py > gpurun_out/r02k_config3_ablation.jsonl 2>gpurun_out/r02k_config3.err
for V in abl1 abl2 abl4 abl8 abl12 abl16 abl32 abl63; do OUZELUM_B200_LIB=$PWD/scratch/variants/lib_$V.so python profiles/time_config3.py >> gpurun_out/r02k_config3_ablation.jsonl 2>>gpurun_out/r02k_config3.err; done
cat gpurun_out/r02k_config3_ablation.jsonl
ncu --set full --clock-control none --import-source on -k regex:ekf_lee_fused -s 70 -c 1 -o gpurun_out/r02k_ekf_fused python profiles/time_config3.py 65536 2 > gpurun_out/r02k_ncu.log 2>&1; tail -2 gpurun_out/r02k_ncu.log
