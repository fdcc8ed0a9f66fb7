#!/bin/bash
# round-2 GPU pass 43: the other one-launch task steps (Landing, LeeLanded, Lando) -- timing and per-instruction captures
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r02as_tasks.jsonl; : > $O
for T in Ouzelum Lando Landing Landed LeeLanded; do python profiles/prof_task.py $T 65536 >> $O 2>>gpurun_out/r02as.err; done
cat $O
ncu --set full --clock-control none --cache-control none --import-source on -k regex:quad_step_kernel -s 70 -c 1 -o gpurun_out/r02as_lee_landed python profiles/prof_task.py LeeLanded 65536 4 > gpurun_out/r02as_ncu1.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:quad_step_kernel -s 70 -c 1 -o gpurun_out/r02as_landing python profiles/prof_task.py Landing 65536 4 > gpurun_out/r02as_ncu2.log 2>&1
