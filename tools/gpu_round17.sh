#!/bin/bash
# LiDAR branch on a second stream: tests + A/B bench
mkdir -p gpurun_out
rm -f gpurun_out/r17_*
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/r17_pytest.log 2>&1
tail -4 gpurun_out/r17_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r17_bench_overlap.json 2> gpurun_out/r17_bench_overlap.err
tail -c 400 gpurun_out/r17_bench_overlap.json; tail -3 gpurun_out/r17_bench_overlap.err
TFPP_NO_OVERLAP=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r17_bench_serial.json 2> gpurun_out/r17_bench_serial.err
tail -c 400 gpurun_out/r17_bench_serial.json; tail -3 gpurun_out/r17_bench_serial.err
