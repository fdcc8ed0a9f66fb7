#!/bin/bash
# round-2 GPU pass 18: steady-state (caches not flushed) vs cold DRAM traffic of the one-launch EKFLeeLanded step
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_bytes.sum
ncu --cache-control none --clock-control none --metrics $M -k regex:ekf_lee_fused -s 40 -c 4 --csv --log-file gpurun_out/r02q_config3_warm_traffic.csv python profiles/time_config3.py 65536 2 > gpurun_out/r02q_ncu1.log 2>&1
ncu --cache-control all --clock-control none --metrics $M -k regex:ekf_lee_fused -s 40 -c 2 --csv --log-file gpurun_out/r02q_config3_cold_traffic.csv python profiles/time_config3.py 65536 2 > gpurun_out/r02q_ncu2.log 2>&1
tail -3 gpurun_out/r02q_ncu1.log
