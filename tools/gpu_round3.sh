#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r3_*.log
for t in test_linear_shapes test_linear_residual_rowmap_stats test_conv3x3 test_grouped_conv test_stem_bn_se test_smallc_conv3x3 test_smallc_wgrad_and_dgrad test_wgrad_dense; do
  timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "$t" 2>&1 | tail -6 >> gpurun_out/r3_ops.log
done
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -x -s 2>&1 | grep -E "got|ratio|passed|failed|Error|assert" | tail -60 > gpurun_out/r3_train.log
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -k "test_forward_eval_vs_golden or test_forward_train_mode" 2>&1 | tail -4 >> gpurun_out/r3_ops.log
timeout 600 python tools/bench_gemm.py > gpurun_out/r3_bench_gemm.log 2>&1
timeout 1500 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/r3_launches.csv env TFPP_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r3_ncu.log 2>&1
cat gpurun_out/r3_ops.log; tail -c 1800 gpurun_out/r3_bench.json; tail -5 gpurun_out/r3_bench.err
