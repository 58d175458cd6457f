#!/bin/bash
# CTA-pair (cta_group::2) conv_gemm: parity with the pair path forced on every eligible shape, then timings / bench A/B.
TAG=${1:-pair}
mkdir -p gpurun_out
TFPP_GEMM_PAIR=2 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_bwd_ops_gpu.py -q -m gpu -x > gpurun_out/${TAG}_tests_forced.log 2>&1; tail -6 gpurun_out/${TAG}_tests_forced.log
for m in 0 1; do for c in qkv; do echo -n "pair=$m "; TFPP_GEMM_PAIR=$m timeout 120 python tools/one_gemm.py $c time 2>&1 | tail -1; done; done > gpurun_out/${TAG}_micro.txt 2>&1; cat gpurun_out/${TAG}_micro.txt
bash tools/gpu_call.sh ${TAG}b bench:off:TFPP_GEMM_PAIR=0 bench
