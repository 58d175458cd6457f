#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/train.log
for t in test_losses_and_gradients_vs_reference test_adamw_step_matches_torch; do
  echo "=== $t" >> gpurun_out/train.log
  timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -k "$t" -x -s 2>&1 | grep -v "^$" | tail -70 >> gpurun_out/train.log
done
for t in test_forward_eval_vs_golden test_blocks_in_isolation test_forward_train_mode_vs_golden; do
  echo "=== $t" >> gpurun_out/train.log
  timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -k "$t" -x -s 2>&1 | grep -v "^$" | grep -E "===|tap |out |worst|passed|failed|Error|assert" | tail -60 >> gpurun_out/train.log
done
tail -5 gpurun_out/train.log
