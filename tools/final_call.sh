#!/bin/bash
# Round-end evidence on the final tree: full GPU suite, smoke, the default bench invocation, the ncu launch list of one
# eager step, the bev_encoder bench leg.
TAG=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -6 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
TFPP_GEMM_DUMP=gpurun_out/${TAG}_gemm_shapes.txt timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<PY
import json
d = json.load(open('gpurun_out/${TAG}_bench.json'))
print('bench', round(d['value'], 1), 'samples/s', round(d['ms_per_step'], 2), 'ms e2e', round(d['e2e']['value'], 1), 'roof', d['roofline']['frac'], 'cpu', d.get('cpu_baseline'), 'launches', d['gpu_launches'], d['clocks'])
print('inference', d['inference'])
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv env TFPP_NO_GRAPH=1 TFPP_PROFILE_STEP=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1; tail -1 gpurun_out/${TAG}_ncu.log | cut -c1-300
TFPP_BENCH_BACKBONE=bev_encoder timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_bev.json 2> gpurun_out/${TAG}_bench_bev.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_bev.json')); print('bev bench', round(d['value'],1), round(d['ms_per_step'],2))"
