"""TEST INFRASTRUCTURE (oracle): CPU fp32 restatement of the ``bev_encoder`` backbone.

Follows team_code/bev_encoder.py:146-233 (``BevEncoder.forward``), :243-272 (``UpsamplingConcat``), :126-137
(``bev_compressor``) and team_code/transfuser_utils.py:596-665 (``create_projection_grid``), as functions of a
state_dict.  RegNet blocks / FPN / heads come from oracle/tfpp_oracle.py.  Pinned to goldens produced by the unmodified
reference (tests/golden/make_golden_bev.py -> bev_b2.npz).

Only tests/, smoke() and bench.py's cpu_baseline may import this module."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import tfpp_oracle as orc

BEV_CFG = dict(orc.DEFAULT_CFG, backbone='bev_encoder', perspective_scale_0=4, perspective_scale_1=2,
               # config.py:100-106,141-142,472: pinhole camera + voxel grid of the lift
               camera_pos=(-1.5, 0.0, 2.0), camera_fov=110, camera_width=1024, camera_height=256, min_z_projection=-10,
               max_z_projection=14, bev_grid_height_downsample_factor=1.0)


def create_projection_grid(cfg):
  """transfuser_utils.py:596-665: (1, D, W, Hh, 3) normalised pixel coordinates (x, y, 0) of every voxel centre of the
  BEV volume + (1, D, W, Hh) visibility mask."""
  mpp = 1.0 / cfg['pixels_per_meter']
  widths = torch.arange(cfg['min_x'], cfg['max_x'], mpp) + (mpp * 0.5)
  depths = torch.arange(cfg['min_y'], cfg['max_y'], mpp) + (mpp * 0.5)
  mpph = mpp * cfg['bev_grid_height_downsample_factor']
  heights = torch.arange(cfg['min_z_projection'], cfg['max_z_projection'], mpph) + (mpph * 0.5)
  depths, widths, heights = torch.meshgrid(depths, widths, heights, indexing='ij')
  cloud = torch.stack((depths, widths, heights), dim=0)
  _, d, w, h = cloud.shape
  t = torch.tensor(cfg['camera_pos']).unsqueeze(1)
  c2 = cloud.view(3, -1) - t                           # identity camera rotation (transfuser_utils.py:620-623)
  c2 = torch.stack((c2[1], c2[2], c2[0]))              # CARLA (x front, y right, z up) -> pinhole (x right, y down, z front)
  f = cfg['camera_width'] / (2.0 * np.tan(cfg['camera_fov'] * np.pi / 360.0))
  k = torch.from_numpy(np.array([[f, 0.0, cfg['camera_width'] / 2.0], [0.0, f, cfg['camera_height'] / 2.0],
                                 [0.0, 0.0, 1.0]])).to(dtype=torch.float32)
  c2 = k @ c2
  z = c2[2:3]
  grid = torch.zeros_like(c2)
  grid[:2] = c2[:2] / z
  grid = grid.view(3, d, w, h)
  ok = (grid[0:1] >= 0.0) & (grid[0:1] < cfg['camera_width']) & (grid[1:2] >= 0.0) & (grid[1:2] < cfg['camera_height']) & \
      (z.view(1, d, w, h) > 0.0)
  grid[0:1] = (grid[0:1] / (0.5 * cfg['camera_width'] - 0.5)) - 1.0
  grid[1:2] = (grid[1:2] / (0.5 * cfg['camera_height'] - 0.5)) - 1.0
  grid = torch.transpose(torch.reshape(grid, [1, 3, d, w, h, 1]), 1, 5).squeeze(1)
  return grid, ok.to(dtype=torch.float32)


def instance_norm(x, eps=1e-5):
  """nn.InstanceNorm2d(affine=False, track_running_stats=False)."""
  mean = x.mean((2, 3), keepdim=True)
  var = x.var((2, 3), unbiased=False, keepdim=True)
  return (x - mean) * torch.rsqrt(var + eps)


def bev_backbone_forward(sd, image, lidar, cfg, training=False, p='backbone', taps=None):
  """BevEncoder.forward (bev_encoder.py:146-233), single-frame LiDAR, RegNet branches.
  Returns (bev feature grid, fused BEV features of stage 3, image features in perspective view)."""
  x = orc.normalize_imagenet(image)
  x = orc._conv_bn(sd, p + '.image_encoder.stem', x, training, stride=2)  # pylint: disable=protected-access
  x = orc.regnet_stage(sd, p + '.image_encoder.s1', x, training)
  x2 = orc.regnet_stage(sd, p + '.image_encoder.s2', x, training)
  x3 = orc.regnet_stage(sd, p + '.image_encoder.s3', x2, training)
  # UpsamplingConcat (bev_encoder.py:243-272)
  up = F.interpolate(x3, size=(x2.shape[2], x2.shape[3]), mode='bilinear', align_corners=False)
  u = torch.cat([x2, up], dim=1)
  u = F.relu(instance_norm(F.conv2d(u, sd[p + '.upsampling_layer.conv.0.weight'], None, padding=1)))
  u = F.relu(instance_norm(F.conv2d(u, sd[p + '.upsampling_layer.conv.3.weight'], None, padding=1)))
  img_feat = F.conv2d(u, sd[p + '.depth_layer.weight'], sd[p + '.depth_layer.bias'])
  if taps is not None:
    taps['image_features'] = img_feat
  # lift (bev_encoder.py:179-199): sample every voxel centre, sum over height, normalise, transpose, mask
  b = lidar.shape[0]
  grid = sd[p + '.grid'].repeat(b, 1, 1, 1, 1)
  vol = F.grid_sample(img_feat.unsqueeze(2), grid, align_corners=False, padding_mode='zeros')
  bev = torch.sum(vol, dim=4) / sd[p + '.bev_projection_normalizer']
  bev = torch.transpose(bev, 2, 3) * sd[p + '.valid_bev_pixels']
  if taps is not None:
    taps['bev_lift'] = bev
  bev = F.gelu(instance_norm(F.conv2d(bev, sd[p + '.bev_compressor.0.weight'], None, padding=1)))
  if taps is not None:
    taps['bev_compressed'] = bev
  f = torch.cat((bev, lidar), dim=1)
  f = orc._conv_bn(sd, p + '.bev_encoder.stem', f, training, stride=2)  # pylint: disable=protected-access
  for i in (1, 2, 3):
    f = orc.regnet_stage(sd, p + f'.bev_encoder.s{i}', f, training)
    if taps is not None:
      taps[f'bev_s{i}'] = f
  p5 = F.relu(F.conv2d(f, sd[p + '.c5_conv.weight'], sd[p + '.c5_conv.bias']))
  p4 = F.interpolate(p5, scale_factor=cfg['bev_upsample_factor'], mode='bilinear', align_corners=False)
  p4 = F.relu(F.conv2d(p4, sd[p + '.up_conv5.weight'], sd[p + '.up_conv5.bias'], padding=1))
  size = (cfg['lidar_resolution_height'] // cfg['bev_down_sample_factor'],
          cfg['lidar_resolution_width'] // cfg['bev_down_sample_factor'])
  p3 = F.interpolate(p4, size=size, mode='bilinear', align_corners=False)
  p3 = F.relu(F.conv2d(p3, sd[p + '.up_conv4.weight'], sd[p + '.up_conv4.bias'], padding=1))
  return p3, f, img_feat


def forward(sd, rgb, lidar_bev, target_point, ego_vel, command, cfg=None, training=False, taps=None):
  """LidarCenterNet.forward (model.py:279-392) with config.backbone = 'bev_encoder'."""
  cfg = cfg or BEV_CFG
  sd = {k: v.float() if torch.is_floating_point(v) else v for k, v in sd.items()}
  feats, fused, grid = bev_backbone_forward(sd, rgb, lidar_bev, cfg, training, taps=taps)
  if taps is not None:
    taps['bev_feature_grid'], taps['fused_features'] = feats, fused
  pl = orc.planner(sd, fused, target_point, ego_vel, command, cfg, training, taps)
  pred_semantic = orc.perspective_decoder(sd, 'semantic_decoder', grid, cfg)
  pred_depth = torch.sigmoid(orc.perspective_decoder(sd, 'depth_decoder', grid, cfg)).squeeze(1)
  bsem = F.relu(F.conv2d(feats, sd['bev_semantic_decoder.0.weight'], sd['bev_semantic_decoder.0.bias'], padding=1))
  bsem = F.conv2d(bsem, sd['bev_semantic_decoder.2.weight'], sd['bev_semantic_decoder.2.bias'])
  bsem = F.interpolate(bsem, size=(cfg['lidar_resolution_height'], cfg['lidar_resolution_width']), mode='bilinear',
                       align_corners=False)
  pred_bev_semantic = bsem * sd['valid_bev_pixels']
  pred_bounding_box = orc.center_net_head(sd, 'head', feats)
  return (None, pl[1], pl[0], pred_semantic, pred_bev_semantic, pred_depth, pred_bounding_box, None, None, None)
