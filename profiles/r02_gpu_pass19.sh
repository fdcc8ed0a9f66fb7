#!/bin/bash
# round-2 GPU pass 19: config 3 -- estimate / command outputs off, L2 evict_last hint on the covariance tile copies
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r02r_config3.jsonl; : > $O
for i in 1 2; do
OZL_EXPOSE_EST=1 timeout 300 python profiles/time_config3.py >> $O 2>gpurun_out/r02r_config3.err
timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02r_config3.err
OUZELUM_B200_LIB=$PWD/scratch/variants/lib_l2hint.so timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02r_config3.err
OUZELUM_B200_LIB=$PWD/scratch/variants/lib_l2hint_coop0.so timeout 300 python profiles/time_config3.py >> $O 2>>gpurun_out/r02r_config3.err
done
cat $O
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_bytes.sum
OUZELUM_B200_LIB=$PWD/scratch/variants/lib_l2hint.so ncu --cache-control none --clock-control none --metrics $M -k regex:ekf_lee_fused -s 40 -c 3 --csv --log-file gpurun_out/r02r_config3_l2hint_traffic.csv python profiles/time_config3.py 65536 2 > gpurun_out/r02r_ncu1.log 2>&1
ncu --cache-control none --clock-control none --metrics $M -k regex:ekf_lee_fused -s 40 -c 3 --csv --log-file gpurun_out/r02r_config3_traffic.csv python profiles/time_config3.py 65536 2 > gpurun_out/r02r_ncu2.log 2>&1
