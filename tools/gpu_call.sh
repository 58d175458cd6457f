#!/bin/bash
# One parameterised GPU-call script (replaces the per-call gpu_round*.sh of round 1).
#   tools/gpu_call.sh <tag> <leg> [<leg> ...]
# legs: tests | ops | model | bench | bench:<name>:<ENV=V,...> | launches | expt
TAG=$1; shift
mkdir -p gpurun_out
rm -f gpurun_out/${TAG}_*
bench_line() {
  python - "$1" <<'PY'
import json, sys
try:
  d = json.load(open(sys.argv[1]))
  r = d.get('roofline') or {}
  print(sys.argv[1], round(d['value'], 1), 'samples/s', round(d['ms_per_step'], 2), 'ms  e2e', round(d['e2e']['value'], 1),
        ' fwd', round(d['inference']['fwd_ms_per_frame'], 2), 'ms/frame  roof', round(r.get('frac', 0), 3), d.get('clocks'))
except Exception as e:
  print(sys.argv[1], 'FAILED', e)
PY
}
for leg in "$@"; do
  case $leg in
    tests) timeout 1800 python -m pytest tests/ -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -8 gpurun_out/${TAG}_pytest.log;;
    ops) timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu > gpurun_out/${TAG}_ops.log 2>&1; tail -8 gpurun_out/${TAG}_ops.log;;
    model) timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_train_gpu.py -q -m gpu > gpurun_out/${TAG}_model.log 2>&1; tail -8 gpurun_out/${TAG}_model.log;;
    bench) TFPP_GEMM_DUMP=gpurun_out/${TAG}_gemm_shapes.txt timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; bench_line gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err;;
    bench:*) IFS=: read -r _ name envs <<< "$leg"; timeout 400 env ${envs//,/ } python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$name.json 2> gpurun_out/${TAG}_bench_$name.err; bench_line gpurun_out/${TAG}_bench_$name.json; tail -2 gpurun_out/${TAG}_bench_$name.err;;
    launches) timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv env TFPP_NO_GRAPH=1 TFPP_PROFILE_STEP=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_ncu.log;;
    expt) timeout 120 env TFPP_EXPERIMENTAL=1 python -m pytest tests/test_ops_gpu.py -q -k "halo_umma" > gpurun_out/${TAG}_halo_tests.log 2>&1; tail -6 gpurun_out/${TAG}_halo_tests.log
          timeout 120 env TFPP_BN_BWD_FUSED=1 python -m pytest tests/test_ops_gpu.py -q -k "bn_backward" > gpurun_out/${TAG}_bnfused_tests.log 2>&1; tail -4 gpurun_out/${TAG}_bnfused_tests.log;;
    *) echo "unknown leg $leg";;
  esac
done
