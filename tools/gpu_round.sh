#!/bin/bash
# one GPU call: block isolation, model parity, new op tests
mkdir -p gpurun_out
timeout 600 python tools/debug_blocks.py > gpurun_out/debug_blocks.log 2>&1
for t in test_forward_eval_vs_golden test_forward_eval_vs_live_oracle test_forward_train_mode_vs_golden test_backbone_module_api; do
  echo "=== $t" >> gpurun_out/model.log
  timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -k "$t" -x -s 2>&1 | grep -v "^$" | tail -45 >> gpurun_out/model.log
done
for t in test_wgrad_dense test_wgrad_grouped_and_stride2; do
  echo "=== $t" >> gpurun_out/ops2.log
  timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "$t" 2>&1 | tail -30 >> gpurun_out/ops2.log
done
tail -5 gpurun_out/ops2.log
