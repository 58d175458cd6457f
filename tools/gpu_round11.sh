#!/bin/bash
# 8-warp epilogue + fp32 fast path, vector wgrad reductions, new attention kernels: full tests, micro-bench, bench, launch list
mkdir -p gpurun_out
rm -f gpurun_out/r11_*
timeout 1500 python -m pytest tests/ -q -m gpu > gpurun_out/r11_pytest.log 2>&1
tail -15 gpurun_out/r11_pytest.log
TFPP_GEMM_DUMP=gpurun_out/r11_gemm_shapes.txt timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r11_bench_n1.json 2> gpurun_out/r11_bench_n1.err
tail -c 1200 gpurun_out/r11_bench_n1.json; tail -3 gpurun_out/r11_bench_n1.err
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r11_launches.csv env TFPP_NO_GRAPH=1 TFPP_PROFILE_STEP=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r11_ncu.log 2>&1
tail -2 gpurun_out/r11_ncu.log
