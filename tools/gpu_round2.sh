#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2_*.log
for t in test_linear_shapes test_linear_residual_rowmap_stats test_conv3x3 test_grouped_conv test_stem_bn_se; do
  timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "$t" 2>&1 | tail -4 >> gpurun_out/r2_ops.log
done
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -x -s 2>&1 | grep -E "got|ratio|passed|failed|Error|assert" | tail -60 > gpurun_out/r2_train.log
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -k "test_forward_eval_vs_golden or test_forward_train_mode" 2>&1 | tail -4 >> gpurun_out/r2_ops.log
timeout 600 python tools/bench_gemm.py > gpurun_out/r2_bench_gemm.log 2>&1
timeout 1500 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
cat gpurun_out/r2_ops.log; tail -c 2500 gpurun_out/r2_bench.json; tail -5 gpurun_out/r2_bench.err
