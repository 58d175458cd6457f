#!/bin/bash
# lean epilogue instantiations of conv_gemm: parity (also with CTA pairs forced), per-shape timing, bench A/B
TAG=${1:-lean}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_bwd_ops_gpu.py -q -m gpu -x > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
TFPP_GEMM_PAIR=2 timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "conv or gemm or linear" > gpurun_out/${TAG}_tests_pair.log 2>&1; tail -2 gpurun_out/${TAG}_tests_pair.log
for m in 0 1; do for c in c576 mlp; do echo -n "lean=$m "; TFPP_GEMM_LEAN=$m timeout 120 python tools/one_gemm.py $c time 2>&1 | tail -1; done; done > gpurun_out/${TAG}_micro.txt 2>&1; cat gpurun_out/${TAG}_micro.txt
bash tools/gpu_call.sh ${TAG}b bench:off:TFPP_GEMM_LEAN=0 bench
