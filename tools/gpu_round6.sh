#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r6_*
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r6_ops.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/r6_launches.csv env TFPP_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r6_ncu.log 2>&1
cat gpurun_out/r6_ops.log
