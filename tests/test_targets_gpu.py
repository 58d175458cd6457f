"""GPU parity of the CenterNet target rasteriser (csrc/targets.cu) with the goldens made by the unmodified reference
(data.py:698-791 via tests/golden/make_targets_golden.py) and with oracle/targets.py on larger random batches."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
KEYS = {'center_heatmap': 'center_heatmap_target', 'wh': 'wh_target', 'offset': 'offset_target',
        'yaw_class': 'yaw_class_target', 'yaw_res': 'yaw_res_target', 'velocity': 'velocity_target',
        'brake': 'brake_target', 'pixel_weight': 'pixel_weight'}


@pytest.fixture(scope='module')
def ops():
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  from carla_garage_b200 import ops as o
  return o


def _pad(cases, nmax):
  boxes = np.zeros((len(cases), nmax, 8), np.float32)
  counts = np.zeros(len(cases), np.int32)
  for i, c in enumerate(cases):
    boxes[i, :len(c)] = c
    counts[i] = len(c)
  return torch.from_numpy(boxes).cuda(), torch.from_numpy(counts).cuda()


def _check(lab, i, want, avg):
  for k, rk in KEYS.items():
    got = lab[k][i].cpu().numpy()
    w = want[rk]
    if got.dtype.kind == 'i':
      assert np.array_equal(got, w.astype(np.int64)), (i, k)
    else:
      assert np.allclose(got, w, rtol=0, atol=2e-6), (i, k, float(np.abs(got - w).max()))
  assert np.array_equal(lab['center_heatmap'][i].cpu().numpy() == 1, want['center_heatmap_target'] == 1)   # exact peaks
  assert float(lab['avg_factor'][i]) == float(avg), i


def test_targets_vs_reference_golden(ops):
  g = np.load(os.path.join(GOLDEN, 'targets.npz'))
  n = sum(1 for k in g.files if k.startswith('boxes'))
  cases = [g[f'boxes{i}'] for i in range(n)]
  boxes, counts = _pad(cases, 32)
  lab = ops.centernet_targets(boxes, counts)
  for i in range(n):
    _check(lab, i, {rk: g[f'{rk}{i}'] for rk in KEYS.values()}, int(g[f'avg{i}']))


def test_targets_vs_oracle_batch_and_loss_path(ops):
  """64 random samples vs oracle/targets.py, and the maps feed the fused loss kernels like host-made labels do."""
  from carla_garage_b200 import synth
  from oracle import targets
  cases = synth.make_gt_boxes(64, seed=17)
  boxes, counts = _pad(cases, 30)
  lab = ops.centernet_targets(boxes, counts)
  for i, c in enumerate(cases):
    want, avg = targets.get_targets(c)
    _check(lab, i, want, avg)
  assert lab['yaw_class'].dtype == torch.int64 and lab['avg_factor'].shape == (64,)
  # same call twice: the kernel clears its outputs itself
  lab2 = ops.centernet_targets(boxes, counts)
  assert all(torch.equal(lab[k], lab2[k]) for k in lab)
