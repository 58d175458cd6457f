#!/bin/bash
mkdir -p gpurun_out
for t in test_forward_eval_vs_golden test_forward_eval_vs_live_oracle test_forward_train_mode_vs_golden test_backbone_module_api; do
  echo "=== $t" >> gpurun_out/model.log
  timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -k "$t" -x -s 2>&1 | tail -60 >> gpurun_out/model.log
done
tail -5 gpurun_out/model.log
