#!/bin/bash
# round-2 GPU pass 28: vehicle step without IEEE-division slow paths on zero numerators
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02aa_pytest.log 2>&1; tail -4 gpurun_out/r02aa_pytest.log
O=gpurun_out/r02aa_config3.jsonl; : > $O
for i in 1 2 3; do timeout 300 python profiles/time_config3.py >> $O 2>gpurun_out/r02aa.err; done
cat $O
