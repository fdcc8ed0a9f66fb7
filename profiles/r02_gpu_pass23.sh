#!/bin/bash
# round-2 GPU pass 23: accel = dv * (1/dt) (the division's slow path ran in every warp: zero numerators), full parity suite, config 3
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02v_pytest.log 2>&1; tail -4 gpurun_out/r02v_pytest.log
O=gpurun_out/r02v_config3.jsonl; : > $O
for i in 1 2 3; do timeout 300 python profiles/time_config3.py >> $O 2>gpurun_out/r02v.err; done
cat $O
