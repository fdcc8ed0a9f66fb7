#!/bin/bash
# round-2 GPU pass 31: noise lambdas (kernel vs oracle, VecTask hooks, graph capture) + full parity suite
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_noise_lambda.py -m gpu -q -x > gpurun_out/r02ad_pytest_noise.log 2>&1; tail -15 gpurun_out/r02ad_pytest_noise.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02ad_pytest.log 2>&1; tail -4 gpurun_out/r02ad_pytest.log
