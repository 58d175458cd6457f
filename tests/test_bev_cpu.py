"""CPU checks of the ``bev_encoder`` backbone: the oracle restatement (oracle/bev_oracle.py) and the
package's host-side geometry against goldens produced by the unmodified reference (tests/golden/make_golden_bev.py)."""
import json
import os

import numpy as np
import torch
import torch.nn.functional as F

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def rel(a, b):
  a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
  return float((a - b).norm() / (b.norm() + 1e-30))


def test_projection_geometry_matches_reference_fingerprints():
  """nn.bev_encoder.projection_grid (closed form) == transfuser_utils.create_projection_grid of the reference: sampled
  grid values, the grid sum, the normaliser and every bit of the visibility mask; the oracle restatement likewise."""
  from carla_garage_b200.config import GlobalConfig
  from carla_garage_b200.nn.bev_encoder import projection_grid
  from oracle import bev_oracle as bo
  g = np.load(os.path.join(GOLDEN, 'bev_b2.npz'))
  for grid, ok in (projection_grid(GlobalConfig()), bo.create_projection_grid(bo.BEV_CFG)):
    assert list(grid.shape) == list(g['grid_shape'])
    assert float(np.abs(grid.flatten()[::997].numpy() - g['grid_sample']).max()) <= 1e-6
    assert abs(float(grid.double().sum()) - float(g['grid_sum'])) <= 1e-3
    norm = torch.finfo(torch.float32).eps + ok.sum(3).unsqueeze(1)
    assert np.array_equal(norm.numpy(), g['normalizer'])
    vbp = torch.transpose(ok.max(3)[0].unsqueeze(1), 2, 3).contiguous()
    assert np.array_equal(np.packbits(vbp.numpy().astype(np.uint8)), g['backbone_valid_bev_pixels'])


def test_lift_tables_reproduce_grid_sample():
  """The separable form the lift kernel evaluates (a_rows, x0, wl, wr) == F.grid_sample over the 256x256x96 volume + sum
  over height + normaliser + transpose + mask (bev_encoder.py:185-199) on a random feature map."""
  from carla_garage_b200.config import GlobalConfig
  from carla_garage_b200.nn.bev_encoder import lift_tables, projection_grid
  grid, ok = projection_grid(GlobalConfig())
  norm = torch.finfo(torch.float32).eps + ok.sum(3).unsqueeze(1)
  vbp = torch.transpose(ok.max(3)[0].unsqueeze(1), 2, 3).contiguous()
  a, x0, wl, wr = lift_tables(grid, norm, vbp, 32, 128)
  assert int(x0.min()) >= 0 and int(x0.max()) <= 126
  torch.manual_seed(0)
  img = torch.randn(2, 3, 32, 128)
  vol = F.grid_sample(img.unsqueeze(2), grid.repeat(2, 1, 1, 1, 1), align_corners=False, padding_mode='zeros')
  want = (vol.sum(4) / norm).transpose(2, 3) * vbp                     # (B, C, W, D)
  v = torch.einsum('dy,bcyx->bcdx', a, img)
  idx = x0.long().view(1, 1, 256, 256).expand(2, 3, -1, -1)
  got = (wl * torch.gather(v, 3, idx) + wr * torch.gather(v, 3, idx + 1)).transpose(2, 3)
  assert rel(got, want) < 1e-6


def test_state_dict_keys_and_oracle_forward_vs_reference_golden():
  """LidarCenterNet(backbone='bev_encoder') has the reference's 1128 state_dict keys in the reference's order and loads
  the seeded state strictly; the oracle restatement reproduces the unmodified reference's eval forward (taps and all
  outputs) on that state."""
  from carla_garage_b200 import synth
  from carla_garage_b200.config import GlobalConfig
  from carla_garage_b200.nn import LidarCenterNet
  from oracle import bev_oracle as bo
  g = np.load(os.path.join(GOLDEN, 'bev_b2.npz'))
  keys = json.load(open(os.path.join(GOLDEN, 'bev_keys.json')))
  cfg = GlobalConfig()
  cfg.backbone = 'bev_encoder'
  m = LidarCenterNet(cfg)
  sd = synth.bev_state(GOLDEN)
  assert list(m.state_dict().keys()) == keys['order']
  assert {k: list(v.shape) for k, v in m.state_dict().items()} == keys['shapes']
  m.load_state_dict(sd, strict=True)
  torch.set_num_threads(min(os.cpu_count() or 1, 16))
  inp = synth.make_inputs(2, seed=11)
  taps = {}
  with torch.no_grad():
    out = bo.forward(sd, inp['rgb'], inp['lidar_bev'], inp['target_point'], inp['ego_vel'], inp['command'], taps=taps)
  errs = {'pred_checkpoint': rel(out[2], g['eval_pred_checkpoint']), 'pred_target_speed': rel(out[1], g['eval_pred_target_speed']),
          'pred_semantic': rel(out[3][..., ::4, ::4], g['eval_pred_semantic']),
          'pred_bev_semantic': rel(out[4][..., ::4, ::4], g['eval_pred_bev_semantic']),
          'pred_depth': rel(out[5][..., ::4, ::4], g['eval_pred_depth'])}
  for n, o in zip(('heatmap', 'wh', 'offset', 'yaw_class', 'yaw_res'), out[6][:5]):
    errs['box_' + n] = rel(o, g['eval_box_' + n])
  for k in ('image_features', 'bev_compressed', 'bev_s1', 'bev_s3'):
    errs['tap_' + k] = rel(taps[k][:, :8, ::2, ::2], g['eval_tap_' + k])
  for k, v in errs.items():
    assert v < 1e-5, (k, v)
