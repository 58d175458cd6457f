import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
  config.addinivalue_line('markers', 'reference: needs /root/reference (build container only)')


@pytest.fixture(scope='session')
def golden_dir():
  return GOLDEN


@pytest.fixture(scope='session')
def state_shapes():
  return json.load(open(os.path.join(GOLDEN, 'state_dict_keys.json')))['shapes']


@pytest.fixture(scope='session')
def oracle_state(state_shapes):
  """The seeded state_dict every golden forward was made with (tests/golden/make_golden.py)."""
  import numpy as np
  import torch
  from carla_garage_b200 import synth
  from oracle import tfpp_oracle as orc
  valid = torch.from_numpy(np.load(os.path.join(GOLDEN, 'valid_bev_pixels.npz'))['valid']).float()
  fixed = {
      'valid_bev_pixels': valid,
      'valid_bev_pixels_inv': 1.0 - valid,
      'loss_speed.weight': torch.tensor(orc.DEFAULT_CFG['target_speed_weights']),
      'loss_semantic.weight': torch.ones(7),
      'loss_bev_semantic.weight': torch.ones(11),
  }
  return synth.make_state_dict(state_shapes, seed=0, fixed=fixed)
