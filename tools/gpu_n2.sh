#!/bin/bash
# 2-GPU leg: NMS / boundary tests on GPU 0, the peer-exchange check, then bench at N=2 with both exchange modes.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nms_gpu.py tests/test_boundary_gpu.py "tests/test_bwd_ops_gpu.py::test_perspective_decoder_forward_backward" -q -m gpu > gpurun_out/n2_tests.log 2>&1; tail -4 gpurun_out/n2_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/peer_check.py > gpurun_out/n2_peer_check.log 2>&1; echo "peer_check rc=$?"; grep '^{' gpurun_out/n2_peer_check.log | tail -1; tail -3 gpurun_out/n2_peer_check.log
for mode in peer nccl; do
  TFPP_EXCHANGE=$mode TFPP_BENCH_ENSEMBLE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n2_bench_$mode.json 2> gpurun_out/n2_bench_$mode.err; echo "bench $mode rc=$?"
  python - gpurun_out/n2_bench_$mode.json <<'PY'
import json,sys
try:
  d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],2),'ms e2e',round(d['e2e']['value'],1), d['config'].get('graphs'), d['config'].get('exchange','')[:30])
except Exception as e: print('FAILED',e)
PY
  tail -2 gpurun_out/n2_bench_$mode.err
done
