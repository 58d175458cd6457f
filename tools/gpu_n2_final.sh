#!/bin/bash
# 2-GPU leg on the final tree: the ensemble test (tolerance fix), then the driver's N=2 bench invocation (peer exchange).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_nms_gpu.py -q -m gpu > gpurun_out/n2f_tests.log 2>&1; tail -2 gpurun_out/n2f_tests.log
TFPP_BENCH_ENSEMBLE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n2f_bench.json 2> gpurun_out/n2f_bench.err; echo "bench rc=$?"
python - gpurun_out/n2f_bench.json <<'PY'
import json,sys
try:
  d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],2),'ms e2e',round(d['e2e']['value'],1), d['config'].get('graphs'), d['config'].get('exchange','')[:40])
except Exception as e: print('FAILED',e)
PY
tail -2 gpurun_out/n2f_bench.err
