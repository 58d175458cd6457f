#!/bin/bash
# A/B of the streaming BatchNorm kernels (TFPP_BN_STREAM=0 = previous kernels): correctness, per-shape timings, bench.
TAG=${1:-bnab}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_bwd_ops_gpu.py -q -m gpu -k "bn or scale_shift or batchnorm or block or se_" > gpurun_out/${TAG}_tests.log 2>&1; tail -5 gpurun_out/${TAG}_tests.log
for mode in 0 1; do TFPP_BN_STREAM=$mode timeout 300 python tools/bn_micro.py; done > gpurun_out/${TAG}_micro.txt 2>&1
cat gpurun_out/${TAG}_micro.txt
bash tools/gpu_call.sh ${TAG}b bench:old:TFPP_BN_STREAM=0 bench
