#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r19_*
timeout 45 env TFPP_EXPERIMENTAL=1 python -m pytest tests/test_ops_gpu.py -q -x -k halo_umma > gpurun_out/r19_halo.log 2>&1
tail -25 gpurun_out/r19_halo.log
