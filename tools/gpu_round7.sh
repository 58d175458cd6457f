#!/bin/bash
# full GPU test-suite in ONE process (what the driver runs) + 2-GPU bench + 1-GPU bench
mkdir -p gpurun_out
rm -f gpurun_out/r7_*
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r7_pytest.log 2>&1
tail -5 gpurun_out/r7_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r7_bench_n2.json 2> gpurun_out/r7_bench_n2.err
tail -c 1200 gpurun_out/r7_bench_n2.json; tail -5 gpurun_out/r7_bench_n2.err
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r7_bench_n1.json 2> gpurun_out/r7_bench_n1.err
tail -c 600 gpurun_out/r7_bench_n1.json
