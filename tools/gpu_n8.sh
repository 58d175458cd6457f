#!/bin/bash
# N-GPU leg (default 8): peer-exchange check, then bench at B=32 and B=12 (BASELINE.json config 3) with the peer exchange,
# and B=12 with NCCL for the A/B.   tools/gpu_n8.sh [N]
N=${1:-8}
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
PEER_CHECK_MODEL=1 run 29521 tools/peer_check.py > gpurun_out/n${N}_peer_check.log 2>&1; echo "peer_check rc=$?"; grep '^{' gpurun_out/n${N}_peer_check.log | tail -1 | cut -c1-1200
for cfg in "peer 32" "peer 12" "nccl 12"; do
  set -- $cfg
  TFPP_EXCHANGE=$1 TFPP_BENCH_ENSEMBLE=0 run 29522 bench.py --gpus $N --steps 10 --warmup 3 --batch $2 --no-cpu-baseline > gpurun_out/n${N}_bench_$1_b$2.json 2> gpurun_out/n${N}_bench_$1_b$2.err; echo "bench $cfg rc=$?"
  python - gpurun_out/n${N}_bench_$1_b$2.json <<'PY'
import json,sys
try:
  d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); print(sys.argv[1], round(d['value'],1), 'samples/s', round(d['ms_per_step'],2),'ms e2e',round(d['e2e']['value'],1), 'graphs', d['config'].get('graphs'))
except Exception as e: print('FAILED',e)
PY
done
