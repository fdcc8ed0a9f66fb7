#!/bin/bash
# round-2 GPU pass 27: steady-state (caches kept) full capture of the fused kernel
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --cache-control none --import-source on -k regex:ekf_lee_fused -s 70 -c 1 -o gpurun_out/r02z_ekf_fused_warm python profiles/time_config3.py 65536 2 > gpurun_out/r02z_ncu.log 2>&1
