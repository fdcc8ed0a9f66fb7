#!/bin/bash
# round-2 GPU pass 17: steady-state DRAM traffic of back-to-back launches (ncu --cache-control none: caches NOT flushed between launches)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_bytes.sum
ncu --cache-control none --clock-control none --metrics $M -k regex:ekf_lee_fused -s 200 -c 4 --csv --log-file gpurun_out/r02q_config3_warm_traffic.csv python profiles/time_config3.py 65536 2 > gpurun_out/r02q_ncu1.log 2>&1
ncu --cache-control all --clock-control none --metrics $M -k regex:ekf_lee_fused -s 200 -c 2 --csv --log-file gpurun_out/r02q_config3_cold_traffic.csv python profiles/time_config3.py 65536 2 > gpurun_out/r02q_ncu2.log 2>&1
for N in 16384 131072 262144; do
ncu --cache-control none --clock-control none --metrics $M -k regex:quad_step -s 100 -c 4 --csv --log-file gpurun_out/r02q_step_${N}_warm_traffic.csv python profiles/prof_step.py $N > gpurun_out/r02q_ncu3.log 2>&1
done
grep -h "dram__bytes\|gpu__time\|hit_rate" gpurun_out/r02q_*_traffic.csv | cut -d, -f5,13- | head -80
