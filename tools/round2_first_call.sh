#!/bin/bash
# First GPU call of the next round: validate and time everything that was written at the end of round 1 without GPU time.
#   1. opt-in GPU tests of the experimental kernels (grouped tcgen05 conv; fused BatchNorm backward through the existing
#      test_bn_backward cases)
#   2. A/B benches: default | TFPP_HALO_UMMA=1 | +TFPP_HALO_UMMA_EPI8=1 | TFPP_HALO_UMMA_GCONV=1 | TFPP_BN_BWD_FUSED=1
# Each leg is wrapped in a timeout; the kernels use bounded mbarrier spins (trap instead of hang).
mkdir -p gpurun_out
rm -f gpurun_out/n2_*
timeout 120 env TFPP_EXPERIMENTAL=1 python -m pytest tests/test_ops_gpu.py -q -k "halo_umma" > gpurun_out/n2_halo_tests.log 2>&1
tail -6 gpurun_out/n2_halo_tests.log
timeout 120 env TFPP_BN_BWD_FUSED=1 python -m pytest tests/test_ops_gpu.py -q -k "bn_backward" > gpurun_out/n2_bnfused_tests.log 2>&1
tail -4 gpurun_out/n2_bnfused_tests.log
run_bench() {  # name, env...
  local name=$1; shift
  timeout 240 env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n2_bench_$name.json 2> gpurun_out/n2_bench_$name.err
  python - <<PY
import json
try:
  d = json.load(open('gpurun_out/n2_bench_$name.json'))
  print('$name', round(d['value'], 1), 'samples/s', round(d['ms_per_step'], 2), 'ms', 'e2e', round(d['e2e']['value'], 1), 'fwd', round(d['inference']['fwd_ms_per_frame'], 2), 'ms/frame')
except Exception as e:
  print('$name FAILED', e)
PY
}
run_bench default TFPP_DUMMY=0
run_bench halo TFPP_HALO_UMMA=1
run_bench halo_epi8 TFPP_HALO_UMMA=1 TFPP_HALO_UMMA_EPI8=1
run_bench halo_gconv TFPP_HALO_UMMA_GCONV=1
run_bench bnfused TFPP_BN_BWD_FUSED=1
run_bench prefetch TFPP_PREFETCH=1
# parity of the whole step with the experimental paths on
timeout 600 env TFPP_HALO_UMMA=1 python -m pytest tests/test_model_gpu.py tests/test_train_gpu.py -q -x > gpurun_out/n2_model_tests_halo.log 2>&1
tail -4 gpurun_out/n2_model_tests_halo.log
