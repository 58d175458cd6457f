#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r18_*
timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r18_bench.json 2> gpurun_out/r18_bench.err
tail -c 600 gpurun_out/r18_bench.json; tail -3 gpurun_out/r18_bench.err
