"""Seeded synthetic inputs, labels and weights of the TransFuser++ step (SURVEY.md §8d "Synthetic inputs").

There is no dataset and no released checkpoint offline, so every test / bench uses these generators.  They are
pure functions of (shape, seed) on the CPU torch generator, hence identical in the build container (where the
reference itself can be run on them to make golden vectors) and on the GPU box.
"""
import math
import zlib

import numpy as np
import torch

N_POINTS = 60000  # 600 k pts/s / 20 Hz * 2 half-sweeps (team_code/config.py:96-99, sensor_agent.py:381)


def _gen(seed, tag=''):
  g = torch.Generator(device='cpu')
  g.manual_seed((int(seed) * 1000003 + zlib.crc32(tag.encode())) % (2**63 - 1))
  return g


def make_point_clouds(batch, seed=1234, n_points=N_POINTS):
  """(B, N, 3) float32: x,y ~ U(-40,40), z ~ U(-1,4)."""
  g = _gen(seed, 'points')
  xy = torch.rand(batch, n_points, 2, generator=g) * 80.0 - 40.0
  z = torch.rand(batch, n_points, 1, generator=g) * 5.0 - 1.0
  return torch.cat([xy, z], dim=2).contiguous()


def make_inputs(batch, seed=1234, lidar_channels=1):
  """rgb (B,3,256,1024) integer-valued f32, lidar_bev (B,C,256,256) in {0,.2,..,1}, target_point, ego_vel, command.

  ``lidar_bev`` here is drawn directly (sparse multiples of 0.2); ``make_point_clouds`` + the pillar-scatter
  kernel give the real thing where the test covers K1 as well.
  """
  g = _gen(seed, 'inputs')
  rgb = torch.randint(0, 256, (batch, 3, 256, 1024), generator=g).float()
  occ = (torch.rand(batch, lidar_channels, 256, 256, generator=g) < 0.35).float()
  lidar = occ * torch.randint(1, 6, (batch, lidar_channels, 256, 256), generator=g).float() / 5.0
  target_point = torch.randn(batch, 2, generator=g) * 10.0
  ego_vel = torch.rand(batch, 1, generator=g) * 8.0
  command = torch.nn.functional.one_hot(torch.randint(0, 6, (batch,), generator=g), 6).float()
  return dict(rgb=rgb, lidar_bev=lidar, target_point=target_point, ego_vel=ego_vel, command=command)


def make_labels(batch, seed=1234):
  """Label tensors with the shapes/dtypes train.py:693-766 moves to the device.  CenterNet targets are drawn
  the way data.py:698-791 / gaussian_target.py:11-61 produce them (gaussian blobs, peak == 1 at box centres)."""
  g = _gen(seed, 'labels')
  lab = {}
  lab['target_speed'] = torch.randint(0, 4, (batch,), generator=g)
  lab['checkpoint'] = torch.cumsum(torch.rand(batch, 10, 2, generator=g), dim=1)
  lab['semantic'] = torch.randint(0, 7, (batch, 256, 1024), generator=g)
  lab['bev_semantic'] = torch.randint(0, 11, (batch, 256, 256), generator=g)
  lab['depth'] = torch.rand(batch, 256, 1024, generator=g)
  hm = torch.zeros(batch, 4, 64, 64)
  wh = torch.zeros(batch, 2, 64, 64)
  off = torch.zeros(batch, 2, 64, 64)
  ycls = torch.zeros(batch, 64, 64, dtype=torch.long)
  yres = torch.zeros(batch, 1, 64, 64)
  pw = torch.zeros(batch, 2, 64, 64)
  avg = torch.zeros(batch)
  yy, xx = torch.meshgrid(torch.arange(64.0), torch.arange(64.0), indexing='ij')
  for b in range(batch):
    n = int(torch.randint(0, 31, (1,), generator=g))
    avg[b] = max(1, n)
    for _ in range(n):
      cx, cy = (int(v) for v in torch.randint(0, 64, (2,), generator=g))
      cls = int(torch.randint(0, 4, (1,), generator=g))
      sigma = 0.5 + 2.0 * float(torch.rand(1, generator=g))
      blob = torch.exp(-((xx - cx)**2 + (yy - cy)**2) / (2 * sigma * sigma))
      blob[blob < 1e-4] = 0
      hm[b, cls] = torch.maximum(hm[b, cls], blob)
      hm[b, cls, cy, cx] = 1.0
      wh[b, :, cy, cx] = torch.rand(2, generator=g) * 8 + 1
      off[b, :, cy, cx] = torch.rand(2, generator=g)
      ycls[b, cy, cx] = int(torch.randint(0, 12, (1,), generator=g))
      yres[b, 0, cy, cx] = (float(torch.rand(1, generator=g)) - 0.5) * (2 * math.pi / 12)
      pw[b, :, cy, cx] = 1.0
  lab.update(center_heatmap=hm, wh=wh, offset=off, yaw_class=ycls, yaw_res=yres, pixel_weight=pw, avg_factor=avg)
  return lab


def make_state_dict(shapes, seed=0, fixed=None):
  """Deterministic, well-conditioned weights for every key in ``shapes`` (name -> shape).

  All BN affine/running stats, position embeddings and queries are *randomised* (the reference zero/one-inits them,
  which would hide residual-branch and BN bugs, SURVEY.md §7 hard part 4).  ``fixed`` holds tensors copied verbatim
  (``valid_bev_pixels`` masks, loss class weights)."""
  fixed = fixed or {}
  sd = {}
  for name in sorted(shapes):
    shape = tuple(shapes[name])
    if name in fixed:
      sd[name] = fixed[name].clone()
      continue
    g = _gen(seed, name)
    leaf = name.rsplit('.', 1)[-1]
    if leaf == 'num_batches_tracked':
      t = torch.zeros(shape, dtype=torch.long)
    elif leaf == 'running_var':
      t = torch.rand(shape, generator=g) + 0.5
    elif leaf == 'running_mean':
      t = torch.randn(shape, generator=g) * 0.1
    elif leaf == 'pos_emb':
      t = torch.randn(shape, generator=g) * 0.1
    elif leaf in ('checkpoint_query', 'wp_query', 'extra_sensor_pos_embed', 'tp_pos_embed'):
      t = torch.rand(shape, generator=g)
    elif 'gru.' in name:
      t = (torch.rand(shape, generator=g) * 2 - 1) / 8.0
    elif len(shape) >= 2:
      fan_in = int(np.prod(shape[1:]))
      gain = 2.0 if len(shape) == 4 else 1.0
      t = torch.randn(shape, generator=g) * math.sqrt(gain / fan_in)
    elif name.endswith('conv3.bn.weight'):
      # last BatchNorm of a residual branch: timm zero-inits it and trained nets keep it small; ~0.2 keeps the
      # residual branch exercised without the DC build-up that makes a random RegNet amplify rounding noise ~7x/stage
      t = torch.rand(shape, generator=g) * 0.2 + 0.1
    elif leaf == 'weight':  # 1-D weights: BatchNorm / LayerNorm gains
      t = torch.rand(shape, generator=g) * 0.4 + 0.8
    else:  # biases
      t = torch.randn(shape, generator=g) * 0.05
    sd[name] = t
  return sd


def golden_state(golden_dir):
  """The seeded + BatchNorm-calibrated state_dict every golden vector was generated with
  (tests/golden/make_golden.py): make_state_dict(seed 0) + the fixed buffers + running stats from bn_calib.npz."""
  import json
  import os
  shapes = json.load(open(os.path.join(golden_dir, 'state_dict_keys.json')))['shapes']
  valid = torch.from_numpy(np.load(os.path.join(golden_dir, 'valid_bev_pixels.npz'))['valid']).float()
  fixed = {
      'valid_bev_pixels': valid,
      'valid_bev_pixels_inv': 1.0 - valid,
      'loss_speed.weight': torch.tensor([0.866605263873406, 7.4527377240841775, 1.2281629310898465, 0.5269622904065803]),
      'loss_semantic.weight': torch.ones(7),
      'loss_bev_semantic.weight': torch.ones(11),
  }
  sd = make_state_dict(shapes, seed=0, fixed=fixed)
  calib = np.load(os.path.join(golden_dir, 'bn_calib.npz'))
  for k in calib.files:
    sd[k] = torch.from_numpy(calib[k]).clone()
  return sd


def make_gt_boxes(n_samples, seed=1234, max_boxes=30):
  """Ground-truth boxes as data.py:937-1000 hands them to get_targets: per sample an (N, 8) float32 array of
  (x, y, extent_x, extent_y, yaw, speed, brake, class) in BEV pixel coordinates of the 256 x 256 LiDAR image."""
  g = _gen(seed, 'gt_boxes')
  out = []
  for _ in range(n_samples):
    n = int(torch.randint(1, max_boxes + 1, (1,), generator=g))
    b = torch.zeros(n, 8)
    b[:, 0:2] = torch.rand(n, 2, generator=g) * 255.0
    b[:, 2] = torch.rand(n, generator=g) * 14.0 + 1.5      # half extents in pixels (4 px / m)
    b[:, 3] = torch.rand(n, generator=g) * 6.0 + 1.5
    b[:, 4] = (torch.rand(n, generator=g) * 2 - 1) * math.pi
    b[:, 5] = torch.rand(n, generator=g) * 10.0
    b[:, 6] = torch.rand(n, generator=g)
    b[:, 7] = torch.randint(0, 4, (n,), generator=g).float()
    out.append(b.numpy().astype(np.float32))
  return out


def make_waypoint_labels(batch, n_wp, seed=1234):
  """ego_waypoints label of train.py:797 (use_wp_gru): (B, pred_len, 2) f32, a forward-moving track."""
  g = _gen(seed, 'waypoints')
  return torch.cumsum(torch.rand(batch, n_wp, 2, generator=g), dim=1)


def mlp_join_state(golden_dir):
  """The seeded state_dict of the transformer_decoder_join = False / use_wp_gru goldens
  (tests/golden/make_golden_mlp_join.py): make_state_dict(seed 0) over that configuration's keys + the fixed buffers."""
  import json
  import os
  shapes = json.load(open(os.path.join(golden_dir, 'mlp_join_keys.json')))['shapes']
  valid = torch.from_numpy(np.load(os.path.join(golden_dir, 'valid_bev_pixels.npz'))['valid']).float()
  fixed = {
      'valid_bev_pixels': valid,
      'valid_bev_pixels_inv': 1.0 - valid,
      'loss_speed.weight': torch.tensor([0.866605263873406, 7.4527377240841775, 1.2281629310898465, 0.5269622904065803]),
      'loss_semantic.weight': torch.ones(7),
      'loss_bev_semantic.weight': torch.ones(11),
  }
  return mlp_join_tweak(make_state_dict(shapes, seed=0, fixed=fixed))


def mlp_join_tweak(sd):
  """The seeded state happens to put the pre-activation of hidden unit 0 of ``join.0`` at 5e-6 (scale 1.2) for sample 0
  of the golden batch: a ReLU kink, where the gradient of ANY two implementations that differ by one rounding is either
  of two very different vectors.  Move that one bias off the kink (generator and tests apply the same shift)."""
  sd['join.0.bias'] = sd['join.0.bias'].clone()
  sd['join.0.bias'][0] += 0.05
  return sd


def bev_state(golden_dir):
  """The seeded state_dict of the ``backbone = 'bev_encoder'`` goldens (tests/golden/make_golden_bev.py):
  make_state_dict(seed 0) over that configuration's keys + the fixed buffers; the three geometry parameters come from
  this package's own closed form (nn.bev_encoder.projection_grid; fingerprints of the reference's are in the golden)."""
  import json
  import os
  from .config import GlobalConfig
  from .nn.bev_encoder import projection_grid
  shapes = json.load(open(os.path.join(golden_dir, 'bev_keys.json')))['shapes']
  valid = torch.from_numpy(np.load(os.path.join(golden_dir, 'valid_bev_pixels.npz'))['valid']).float()
  cfg = GlobalConfig()
  grid, ok = projection_grid(cfg)
  fixed = {
      'valid_bev_pixels': valid,
      'valid_bev_pixels_inv': 1.0 - valid,
      'loss_speed.weight': torch.tensor([0.866605263873406, 7.4527377240841775, 1.2281629310898465, 0.5269622904065803]),
      'loss_semantic.weight': torch.ones(7),
      'loss_bev_semantic.weight': torch.ones(11),
      'backbone.grid': grid,
      'backbone.bev_projection_normalizer': torch.finfo(torch.float32).eps + torch.sum(ok, dim=3).unsqueeze(1),
      'backbone.valid_bev_pixels': torch.transpose(torch.max(ok, dim=3)[0].unsqueeze(1), 2, 3).contiguous(),
  }
  return make_state_dict(shapes, seed=0, fixed=fixed)
