import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


# Golden vectors and oracle comparisons are dropout-free (masks cannot match torch's generator); the dropout path has
# its own tests (tests/test_dropout_gpu.py) that switch it on per engine and feed the same masks to the oracle.
os.environ.setdefault('TFPP_DROPOUT', '0')


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
  config.addinivalue_line('markers', 'reference: needs /root/reference (build container only)')
  # the CPU oracle (torch fp32) collapses when oversubscribed on many-core hosts: 16 threads are ~10x faster than 128
  import torch
  torch.set_num_threads(min(os.cpu_count() or 1, 16))
  # "fp32 torch reference" must mean fp32: cuDNN / cuBLAS would otherwise be free to run the references in TF32
  torch.backends.cudnn.allow_tf32 = False
  torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope='session')
def golden_dir():
  return GOLDEN


@pytest.fixture(scope='session')
def state_shapes():
  return json.load(open(os.path.join(GOLDEN, 'state_dict_keys.json')))['shapes']


@pytest.fixture(scope='session')
def oracle_state():
  """The seeded, BatchNorm-calibrated state_dict every golden forward was made with (tests/golden/make_golden.py)."""
  from carla_garage_b200 import synth
  return synth.golden_state(GOLDEN)
