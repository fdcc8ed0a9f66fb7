#!/bin/bash
# round-2 GPU pass 32: record of the final binary -- parity suite, smoke, bench (both arms at the driver's flags, K = 1000), launch list,
# size sweep, config 3, full ncu captures of the three hot kernels
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02ae_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ae_pytest.log
tail -4 gpurun_out/r02ae_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ae_smoke.log 2>&1; tail -1 gpurun_out/r02ae_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02ae_bench_k20.json 2> gpurun_out/r02ae_bench_k20.err; echo rc=$?
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02ae_bench_ref.json 2>/dev/null; echo rc=$?
python bench.py --steps 1000 --warmup 100 --no-side-configs --no-cpu-baseline > gpurun_out/r02ae_bench_k1000.json 2> gpurun_out/r02ae_bench_k1000.err; echo rc=$?
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02ae_launches.csv python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/r02ae_ncu_launch.log 2>&1
python profiles/time_sizes.py > gpurun_out/r02ae_sizes.json 2>&1
python profiles/time_config3.py > gpurun_out/r02ae_config3.jsonl 2> gpurun_out/r02ae_config3.err
python profiles/time_config3.py 262144 100 >> gpurun_out/r02ae_config3.jsonl 2>> gpurun_out/r02ae_config3.err
ncu --set full --clock-control none --import-source on -k regex:ekf_lee_fused -s 70 -c 1 -o gpurun_out/r02ae_ekf_fused python profiles/time_config3.py 65536 2 > gpurun_out/r02ae_ncu_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:quad_step_tma -s 5 -c 1 -o gpurun_out/r02ae_quad_step_tma_1M python profiles/prof_step.py 1048576 20 > gpurun_out/r02ae_ncu_tma.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:quad_step_kernel -s 80 -c 1 -o gpurun_out/r02ae_quad_step_16k python profiles/prof_step.py 16384 > gpurun_out/r02ae_ncu16k.log 2>&1
ls -la gpurun_out/*.ncu-rep
