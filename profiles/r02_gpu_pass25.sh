#!/bin/bash
# round-2 GPU pass 25: full capture of the fused kernel after the tensor-map / reciprocal changes
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ekf_lee_fused -s 70 -c 1 -o gpurun_out/r02x_ekf_fused python profiles/time_config3.py 65536 2 > gpurun_out/r02x_ncu_fused.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:ekf_lee_fused -s 70 -c 1 -o gpurun_out/r02x_ekf_fused_warm python profiles/time_config3.py 65536 2 > gpurun_out/r02x_ncu_fused2.log 2>&1
ls -la gpurun_out/r02x*
