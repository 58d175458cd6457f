"""BevEncoder with the reference's constructor, attributes and state_dict keys (team_code/bev_encoder.py:15-139,
243-272): SimpleBEV-style lift of the camera features into the BEV grid, concatenated with the LiDAR histogram, one
RegNetY-3.2GF over the fused BEV.  Parameter containers only — the forward pass is
``engine.Engine.bev_backbone_forward`` on the sm_100a kernels (csrc/bev_lift.cu + the shared conv / norm kernels).

Geometry: ``grid`` / ``bev_projection_normalizer`` / ``valid_bev_pixels`` are kept as (frozen) parameters exactly like
the reference does (they are state_dict entries of its checkpoints), but the device never reads the 75 MB grid:
``lift_tables`` folds it into one (depth, image rows) matrix and three (depth, width) tables, see csrc/bev_lift.cu."""
import numpy as np
import torch
from torch import nn

from . import regnet
from .. import engine


def projection_grid(config):
  """The voxel -> normalised-pixel map of transfuser_utils.py:596-665 (create_projection_grid), built from the closed
  form of a pinhole camera without rotation: for the voxel centre (X front, Y right, Z up) in camera-centred metres
      u = f * Y / X + cx   depends on (depth, width) only,     v = f * Z / X + cy   on (depth, height) only
  (the reference feeds +Z into the 'down' axis of its pinhole frame; restated as is).  Returns
  (grid (1, D, W, H, 3) f32 in normalised display coordinates with a zero third component, valid (1, D, W, H) f32)."""
  assert tuple(config.camera_rot_0) == (0.0, 0.0, 0.0), 'a rotated camera is not separable; only the stock mounting is built'
  mpp = 1.0 / config.pixels_per_meter
  mpp_h = mpp * config.bev_grid_height_downsample_factor
  f32 = torch.float32
  ys = torch.arange(config.min_x, config.max_x, mpp, dtype=f32) + mpp * 0.5            # width axis  (CARLA y, right)
  xs = torch.arange(config.min_y, config.max_y, mpp, dtype=f32) + mpp * 0.5            # depth axis  (CARLA x, front)
  zs = torch.arange(config.min_z_projection, config.max_z_projection, mpp_h, dtype=f32) + mpp_h * 0.5
  cam = torch.tensor(config.camera_pos, dtype=f32)
  focal = np.float32(config.camera_width / (2.0 * np.tan(config.camera_fov * np.pi / 360.0)))
  cx, cy = np.float32(config.camera_width / 2.0), np.float32(config.camera_height / 2.0)
  xd = (xs - cam[0]).view(-1, 1)                                 # (D, 1) distance in front of the camera
  # the reference multiplies by the intrinsic matrix first (f * Y + cx * X) and divides by X afterwards
  u = (focal * (ys - cam[1]).view(1, -1) + cx * xd) / xd         # (D, W)
  v = (focal * (zs - cam[2]).view(1, -1) + cy * xd) / xd         # (D, H)
  d, w, h = xs.numel(), ys.numel(), zs.numel()
  ok = ((u >= 0.0) & (u < config.camera_width)).view(d, w, 1) & ((v >= 0.0) & (v < config.camera_height)).view(d, 1, h) & \
      (xd > 0.0).view(d, 1, 1)
  grid = torch.zeros((1, d, w, h, 3), dtype=f32)
  grid[0, :, :, :, 0] = (u / (0.5 * config.camera_width - 0.5) - 1.0).view(d, w, 1)
  grid[0, :, :, :, 1] = (v / (0.5 * config.camera_height - 0.5) - 1.0).view(d, 1, h)
  return grid, ok.to(f32).unsqueeze(0)


def lift_tables(grid, normalizer, valid_bev_pixels, img_h, img_w):
  """Fold F.grid_sample(bilinear, zeros padding, align_corners=False) + the sum over height + the normaliser + the
  visibility mask (bev_encoder.py:185-199) into the tables of tfpp_bev_lift.  grid (1, D, W, H, 3), normalizer
  (1, 1, D, W), valid_bev_pixels (1, 1, W, D) — the module's own parameters (so a loaded checkpoint is honoured).
  Returns (a_rows (D, img_h) f32, x0 (D, W) int32, wl (D, W) f32, wr (D, W) f32) on the CPU."""
  g = grid.detach().double().cpu()[0]
  d, w, h, _ = g.shape
  gx, gy = g[..., 0], g[..., 1]
  if float((gx - gx[:, :, :1]).abs().max()) > 1e-6 or float((gy - gy[:, :1, :]).abs().max()) > 1e-6 or \
      float(g[..., 2].abs().max()) != 0.0:
    raise NotImplementedError('bev lift: the projection grid is not separable (rotated camera?)')
  ix = ((gx[:, :, 0] + 1.0) * img_w - 1.0) / 2.0   # (D, W) un-normalised column, align_corners=False
  iy = ((gy[:, 0, :] + 1.0) * img_h - 1.0) / 2.0   # (D, H) un-normalised row
  # vertical: A[d, y] = sum_h weight of row y
  y0 = torch.floor(iy)
  wy1 = iy - y0
  a_rows = torch.zeros((d, img_h), dtype=torch.float64)
  for yy, ww in ((y0, 1.0 - wy1), (y0 + 1.0, wy1)):
    inside = (yy >= 0) & (yy <= img_h - 1)
    a_rows.scatter_add_(1, yy.clamp(0, img_h - 1).long(), torch.where(inside, ww, torch.zeros_like(ww)))
  # horizontal: two neighbours, re-expressed on a base column in [0, img_w - 2]
  x0 = torch.floor(ix)
  wx1 = ix - x0
  wx0 = 1.0 - wx1
  left_in = (x0 >= 0) & (x0 <= img_w - 1)
  right_in = (x0 + 1 >= 0) & (x0 + 1 <= img_w - 1)
  wl = torch.where(left_in, wx0, torch.zeros_like(wx0))
  wr = torch.where(right_in, wx1, torch.zeros_like(wx1))
  base = x0.clamp(0, img_w - 2)
  shift_r = x0 < 0               # only the right neighbour (column 0) can be inside: it becomes the base column
  shift_l = x0 > img_w - 2       # only the left neighbour (column img_w - 1) can be inside: it becomes base + 1
  wl2 = torch.where(shift_r, torch.where(x0 == -1, wr, torch.zeros_like(wr)), torch.where(shift_l, torch.zeros_like(wl), wl))
  wr2 = torch.where(shift_l, torch.where(x0 == img_w - 1, wl, torch.zeros_like(wl)), torch.where(shift_r, torch.zeros_like(wr), wr))
  scale = valid_bev_pixels.detach().double().cpu()[0, 0].t() / normalizer.detach().double().cpu()[0, 0]  # (D, W)
  return (a_rows.float().contiguous(), base.to(torch.int32).contiguous(), (wl2 * scale).float().contiguous(),
          (wr2 * scale).float().contiguous())


class UpsamplingConcat(regnet._NoForward):  # pylint: disable=protected-access
  """bev_encoder.py:243-272: bilinear up-sampling + concatenation skip + 2 x (conv3x3, InstanceNorm2d, ReLU)."""

  def __init__(self, in_channels, out_channels):
    super().__init__()
    self.conv = nn.Sequential(
        nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1, bias=False),
        nn.InstanceNorm2d(out_channels),
        nn.ReLU(inplace=True),
        nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1, bias=False),
        nn.InstanceNorm2d(out_channels),
        nn.ReLU(inplace=True),
    )


class BevEncoder(nn.Module):
  """Bev sensor fusion backbone (bev_encoder.py:15-139), CUDA-native forward."""

  def __init__(self, config):
    super().__init__()
    self.config = config
    if config.image_architecture != 'regnety_032' or config.lidar_architecture != 'regnety_032':
      raise NotImplementedError('carla_garage_b200 builds the regnety_032 branches only (video backbones: SURVEY.md §8 f4)')
    self.lidar_video = False
    in_channels = (2 if config.use_ground_plane else 1) * config.lidar_seq_len
    # parameters first: the reference registers them on the module itself, so they lead its state_dict
    grid, valid_voxels = projection_grid(config)
    self.image_encoder = regnet.RegNetY032Features(in_chans=3, num_stages=3)
    info = self.image_encoder.feature_info.info
    img_start_index = 1
    self.perspective_upsample_factor = info[img_start_index + 2]['reduction'] // config.perspective_downsample_factor
    self.avgpool_img = nn.AdaptiveAvgPool2d((config.img_vert_anchors, config.img_horz_anchors))
    self.bev_encoder = regnet.RegNetY032Features(in_chans=in_channels + config.bev_latent_dim, num_stages=3)
    self.global_pool_bev = nn.AdaptiveAvgPool2d(output_size=1)
    self.avgpool_lidar = nn.AdaptiveAvgPool2d((config.lidar_vert_anchors, config.lidar_horz_anchors))
    self.global_pool_img = nn.AdaptiveAvgPool2d(output_size=1)
    self.num_features = self.bev_encoder.feature_info.info[img_start_index + 2]['num_chs']
    if config.detect_boxes or config.use_bev_semantic:
      channel = config.bev_features_chanels
      self.relu = nn.ReLU(inplace=True)
      self.upsample = nn.Upsample(scale_factor=config.bev_upsample_factor, mode='bilinear', align_corners=False)
      self.upsample2 = nn.Upsample(size=(config.lidar_resolution_height // config.bev_down_sample_factor,
                                         config.lidar_resolution_width // config.bev_down_sample_factor),
                                   mode='bilinear', align_corners=False)
      self.up_conv5 = nn.Conv2d(channel, channel, (3, 3), padding=1)
      self.up_conv4 = nn.Conv2d(channel, channel, (3, 3), padding=1)
      self.c5_conv = nn.Conv2d(self.num_features, channel, (1, 1))
    self.grid = nn.Parameter(grid, requires_grad=False)
    normalizer = torch.finfo(torch.float32).eps + torch.sum(valid_voxels, dim=3).unsqueeze(1)
    self.bev_projection_normalizer = nn.Parameter(normalizer, requires_grad=False)
    valid_bev_pixels = torch.transpose(torch.max(valid_voxels, dim=3)[0].unsqueeze(1), 2, 3).contiguous()
    self.valid_bev_pixels = nn.Parameter(valid_bev_pixels, requires_grad=False)
    self.upsampling_layer = UpsamplingConcat(info[img_start_index + 1]['num_chs'] + info[img_start_index + 2]['num_chs'],
                                             config.image_u_net_output_features)
    self.depth_layer = nn.Conv2d(config.image_u_net_output_features, config.bev_latent_dim, kernel_size=1, padding=0)
    self.bev_compressor = nn.Sequential(
        nn.Conv2d(config.bev_latent_dim, config.bev_latent_dim, kernel_size=3, padding=1, stride=1, bias=False),
        nn.InstanceNorm2d(config.bev_latent_dim),
        nn.GELU(),
    )
    self.num_image_features = config.bev_latent_dim

  def forward(self, image, lidar):
    """Same contract as bev_encoder.py:146-233: NCHW f32 in, (features, fused_bev_features, image_features) NCHW f32."""
    from .. import ops  # pylint: disable=import-outside-toplevel
    eng = engine.Engine.for_backbone(self)
    feats, fused, grid = eng.bev_backbone_forward(image, lidar, training=self.training)
    fused = fused.float() if fused.dim() == 2 else ops.nhwc_to_nchw(fused)
    return (ops.nhwc_to_nchw(feats) if feats is not None else None, fused, ops.nhwc_to_nchw(grid))
