#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r5_*.log
for t in test_linear_shapes test_linear_residual_rowmap_stats test_conv3x3 test_grouped_conv test_stem_bn_se; do
  timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "$t" 2>&1 | tail -6 >> gpurun_out/r5_ops.log
done
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -x 2>&1 | tail -25 >> gpurun_out/r5_ops.log
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -k "test_forward_eval_vs_golden or test_forward_train_mode" 2>&1 | tail -4 >> gpurun_out/r5_ops.log
timeout 600 python tools/bench_gemm.py > gpurun_out/r5_bench_gemm.log 2>&1
timeout 1500 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err
cat gpurun_out/r5_ops.log; tail -c 1500 gpurun_out/r5_bench.json; tail -5 gpurun_out/r5_bench.err
