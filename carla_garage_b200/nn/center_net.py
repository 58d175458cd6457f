"""LidarCenterNetHead with the reference's constructor / state_dict keys (team_code/center_net.py:12-47); forward,
loss and decode run on the sm_100a kernels."""
from torch import nn

from . import regnet


class LidarCenterNetHead(nn.Module):
  """Objects-as-points head (center_net.py:12-75,142-237)."""

  def __init__(self, config):
    super().__init__()
    self.config = config
    self.heatmap_head = self._build_head(config.bb_input_channel, config.num_bb_classes)
    self.wh_head = self._build_head(config.bb_input_channel, 2)
    self.offset_head = self._build_head(config.bb_input_channel, 2)
    self.yaw_class_head = self._build_head(config.bb_input_channel, config.num_dir_bins)
    self.yaw_res_head = self._build_head(config.bb_input_channel, 1)
    if not (config.lidar_seq_len == 1 and config.seq_len == 1):
      raise NotImplementedError('temporal velocity / brake heads are outside the single-frame TransFuser++ path')

  def _build_head(self, in_channel, out_channel):
    return nn.Sequential(nn.Conv2d(in_channel, in_channel, kernel_size=3, padding=1), nn.ReLU(inplace=True),
                         nn.Conv2d(in_channel, out_channel, kernel_size=1))

  def head_names(self):
    return ('heatmap_head', 'wh_head', 'offset_head', 'yaw_class_head', 'yaw_res_head')

  def forward(self, feat):
    """feat: NCHW f32 (B,64,64,64) like the reference (center_net.py:49-75); returns the 7-tuple."""
    from .. import engine, ops  # pylint: disable=import-outside-toplevel
    eng = engine.Engine.for_head(self)
    return eng.center_head_forward(ops.nchw_to_nhwc(feat.contiguous()))

  def get_bboxes(self, center_heatmap_preds, wh_preds, offset_preds, yaw_class_preds, yaw_res_preds,
                 velocity_preds=None, brake_preds=None):
    """center_net.py:142-170 -> decode_heatmap (center_net.py:172-237) on the GPU: (B, k, 9)."""
    del velocity_preds, brake_preds
    from .. import ops  # pylint: disable=import-outside-toplevel
    return ops.decode_heatmap(center_heatmap_preds, wh_preds, offset_preds, yaw_class_preds, yaw_res_preds,
                              k=self.config.top_k_center_keypoints, img_h=self.config.lidar_resolution_height,
                              img_w=self.config.lidar_resolution_width)

  def loss(self, center_heatmap_pred, wh_pred, offset_pred, yaw_class_pred, yaw_res_pred, velocity_pred, brake_pred,
           center_heatmap_target, wh_target, yaw_class_target, yaw_res_target, offset_target, velocity_target,
           brake_target, pixel_weight, avg_factor):
    """center_net.py:77-123 on tfpp_center_head_loss: dict of the five head losses (gaussian focal heat-map loss, masked
    L1 wh / offset, masked CE yaw class, masked SmoothL1 yaw residual, each / sum(avg_factor)).  Values only: the
    gradient path of the training step runs through LidarCenterNet.compute_loss (carla_garage_b200.boundary), which
    evaluates the same kernel together with the other five losses."""
    del velocity_pred, brake_pred, velocity_target, brake_target
    import torch  # pylint: disable=import-outside-toplevel
    from .. import _lib, ops  # pylint: disable=import-outside-toplevel
    maps = torch.cat([center_heatmap_pred, wh_pred, offset_pred, yaw_class_pred, yaw_res_pred], dim=1).float().contiguous()
    b, _, h, w = maps.shape
    dev = maps.device
    f = lambda t: t.to(dev, torch.float32).contiguous()
    sums = torch.zeros(5, dtype=torch.float32, device=dev)
    w5 = torch.ones(5, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().tfpp_center_head_loss(
        maps.data_ptr(), f(center_heatmap_target).data_ptr(), f(wh_target).data_ptr(), f(offset_target).data_ptr(),
        yaw_class_target.to(dev, torch.long).contiguous().data_ptr(), f(yaw_res_target).data_ptr(),
        f(pixel_weight).data_ptr(), f(avg_factor).data_ptr(), w5.data_ptr(), sums.data_ptr(), None, None, b, h * w,
        self.config.num_bb_classes, self.config.num_dir_bins, 24, ops._stream()), 'center loss')  # pylint: disable=protected-access
    keys = ('loss_center_heatmap', 'loss_wh', 'loss_offset', 'loss_yaw_class', 'loss_yaw_res')
    return {k: sums[i] for i, k in enumerate(keys)}

  decode_heatmap = None  # reference-internal helper; use get_bboxes


del regnet
