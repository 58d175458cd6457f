#!/bin/bash
# run every GPU op test in its own process (a device trap must not poison the following tests)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for t in test_pillar_scatter_bit_exact test_linear_shapes test_linear_residual_rowmap_stats test_conv3x3 test_grouped_conv test_stem_bn_se test_pool_bilinear_layout test_layernorm_attention test_planner_kernels test_decode_heatmap; do
  echo "=== $t" >> gpurun_out/ops.log
  timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "$t" -x 2>&1 | tail -25 >> gpurun_out/ops.log
done
tail -5 gpurun_out/ops.log
