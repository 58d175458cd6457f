#!/bin/bash
# compute-sanitizer evidence: memcheck over one eager training step + the agent kernels, racecheck over
# the shared-memory-heavy kernels one at a time, and the allocator stress run of the multi-stream backward.
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/one_step.py 1 > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitize_memcheck.log
for c in attn576 battn576 gconv smallc nms targets qkv s1stats pillar; do
  timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/one_gemm.py $c > gpurun_out/sanitize_race_$c.log 2>&1; echo "racecheck $c rc=$? $(grep -c 'hazard' gpurun_out/sanitize_race_$c.log) hazard lines"; tail -1 gpurun_out/sanitize_race_$c.log
done
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 600 python tools/one_step.py 3 > gpurun_out/sanitize_nocache.log 2>&1; echo "no-caching-allocator run rc=$?"; tail -3 gpurun_out/sanitize_nocache.log
timeout 300 python tools/one_step.py 3 > gpurun_out/sanitize_cached.log 2>&1; tail -3 gpurun_out/sanitize_cached.log
