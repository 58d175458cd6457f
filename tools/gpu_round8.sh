#!/bin/bash
# new kernels (elementwise rewrites, SE, PackPlan): tests, micro-bench, bench, launch list of one timed eager step
mkdir -p gpurun_out
rm -f gpurun_out/r8_*
timeout 1500 python -m pytest tests/ -q -m gpu > gpurun_out/r8_pytest.log 2>&1
tail -15 gpurun_out/r8_pytest.log
timeout 600 python tools/bench_elem.py > gpurun_out/r8_bench_elem.log 2>&1
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r8_bench_n1.json 2> gpurun_out/r8_bench_n1.err
tail -c 1500 gpurun_out/r8_bench_n1.json; tail -3 gpurun_out/r8_bench_n1.err
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r8_launches.csv env TFPP_NO_GRAPH=1 TFPP_PROFILE_STEP=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r8_ncu.log 2>&1
tail -2 gpurun_out/r8_ncu.log
