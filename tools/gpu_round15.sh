#!/bin/bash
# what the driver runs at round end: GPU suite (-x), smoke(), default bench (with CPU baseline), reference arm
mkdir -p gpurun_out
rm -f gpurun_out/r15_*
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/r15_pytest.log 2>&1
tail -4 gpurun_out/r15_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r15_smoke.log 2>&1
tail -2 gpurun_out/r15_smoke.log
timeout 900 python bench.py > gpurun_out/r15_bench_default.json 2> gpurun_out/r15_bench_default.err
tail -c 1500 gpurun_out/r15_bench_default.json; tail -3 gpurun_out/r15_bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r15_bench_reference.json 2> gpurun_out/r15_bench_reference.err
tail -c 600 gpurun_out/r15_bench_reference.json; tail -3 gpurun_out/r15_bench_reference.err
