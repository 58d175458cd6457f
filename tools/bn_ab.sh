#!/bin/bash
# A/B of the BatchNorm kernel variants (TFPP_BN_STREAM=0 round-1 kernels, 1 streaming (default), 2 cp.async ring + U=8
# forward apply): correctness, per-shape timings, bench.
TAG=${1:-bnab}
A=${2:-1}
B=${3:-2}
mkdir -p gpurun_out
TFPP_BN_STREAM=$B timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_bwd_ops_gpu.py -q -m gpu -k "bn or scale_shift or batchnorm or block or se_" > gpurun_out/${TAG}_tests.log 2>&1; tail -5 gpurun_out/${TAG}_tests.log
for mode in $A $B; do TFPP_BN_STREAM=$mode timeout 300 python tools/bn_micro.py; done > gpurun_out/${TAG}_micro.txt 2>&1
grep -E "sum|i1 |i3 |l3 " gpurun_out/${TAG}_micro.txt
bash tools/gpu_call.sh ${TAG}b bench:a:TFPP_BN_STREAM=$A bench:b:TFPP_BN_STREAM=$B
