#!/bin/bash
mkdir -p gpurun_out
TFPP_BENCH_ENSEMBLE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n4f_bench.json 2> gpurun_out/n4f_bench.err; echo "bench rc=$?"
python - gpurun_out/n4f_bench.json <<'PY'
import json,sys
try:
  d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],2),'ms e2e',round(d['e2e']['value'],1), d['config'].get('graphs'), d['config'].get('exchange','')[:40])
except Exception as e: print('FAILED',e)
PY
tail -2 gpurun_out/n4f_bench.err
