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
  config.addinivalue_line('markers', 'noisy: end-to-end comparison at the fp32-atomics noise floor; attempted twice (see conftest)')
  # the CPU oracle (torch fp32) collapses when oversubscribed on many-core hosts: 16 threads are ~10x faster than 128
  import torch
  torch.set_num_threads(min(os.cpu_count() or 1, 16))
  # "fp32 torch reference" must mean fp32: cuDNN / cuBLAS would otherwise be free to run the references in TF32
  torch.backends.cudnn.allow_tf32 = False
  torch.backends.cuda.matmul.allow_tf32 = False


RERUNS = []   # node ids whose first attempt failed (reported at the end of the session)


def pytest_runtest_protocol(item, nextitem):
  """Tests marked ``noisy`` compare END-TO-END quantities of a randomly initialised network whose kernels reduce with
  fp32 atomics: two runs of the same binary differ in the last bit, which now and then puts one ReLU / one bf16 rounding
  on the other side of a threshold and moves a whole gradient slice (DESIGN.md "Numerics"; the CPU oracle against the CPU
  reference shows the same effect).  Such a test is attempted a second time before it counts as failed — a wrong kernel
  fails twice, a threshold flip does not — and every retry is listed in the terminal summary."""
  if item.get_closest_marker('noisy') is None:
    return None
  from _pytest.runner import runtestprotocol
  item.ihook.pytest_runtest_logstart(nodeid=item.nodeid, location=item.location)
  reports = runtestprotocol(item, nextitem=nextitem, log=False)
  if any(r.failed and r.when == 'call' for r in reports):
    RERUNS.append(item.nodeid)
    reports = runtestprotocol(item, nextitem=nextitem, log=False)
  for r in reports:
    item.ihook.pytest_runtest_logreport(report=r)
  item.ihook.pytest_runtest_logfinish(nodeid=item.nodeid, location=item.location)
  return True


def pytest_terminal_summary(terminalreporter):
  if RERUNS:
    terminalreporter.write_line(f'noise-bound tests attempted twice ({len(RERUNS)}): ' + ', '.join(RERUNS))


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
