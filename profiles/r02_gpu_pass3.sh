#!/bin/bash
# round-2 GPU pass 3: full parity suite on the ABI-v3 build, config-3 kernel variants, bench at the driver's K
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -25 gpurun_out/r02c_pytest.log
for B in 128 256 512; do OZL_EKF_BLOCK=$B python profiles/time_config3.py >> gpurun_out/r02c_config3.jsonl 2>> gpurun_out/r02c_config3.err; done
for B in 128 512; do OUZELUM_B200_LIB=$PWD/scratch/variants/lib_philox_inline.so OZL_EKF_BLOCK=$B python profiles/time_config3.py >> gpurun_out/r02c_config3.jsonl 2>> gpurun_out/r02c_config3.err; done
cat gpurun_out/r02c_config3.jsonl
python profiles/time_sizes.py 16384 131072 262144 1048576 > gpurun_out/r02c_sizes.json 2>&1; cat gpurun_out/r02c_sizes.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench_k20.json 2> gpurun_out/r02c_bench_k20.err; echo rc=$?
