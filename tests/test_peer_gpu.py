"""NVLink peer exchange (csrc/peer_exchange.cu) on >= 2 GPUs: spawns tools/peer_check.py under torchrun.  Skipped on a
one-GPU box (the driver's `pytest -m gpu` run); run with `gpurun --gpus 2`."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_exchange_matches_nccl_allreduce_plus_adamw():
  n = torch.cuda.device_count() if torch.cuda.is_available() else 0
  if n < 2:
    pytest.skip('needs >= 2 GPUs')
  n = 2 if n < 4 else (4 if n < 8 else 8)
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr', '127.0.0.1',
         '--master-port', '29517', os.path.join(ROOT, 'tools', 'peer_check.py')]
  out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
  lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
  assert out.returncode == 0 and lines, out.stdout[-2000:] + out.stderr[-2000:]
  res = json.loads(lines[-1])
  assert res['ok'] and res['model']['graphs'] == 1
