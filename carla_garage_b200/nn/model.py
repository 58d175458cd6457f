"""LidarCenterNet with the reference's constructor, forward signature, attributes and state_dict keys
(team_code/model.py:24-445), running on the sm_100a kernels through carla_garage_b200.engine.

Built configuration = the TransFuser++ default (``GlobalConfig()``): transFuser backbone, transformer-decoder join,
checkpoint + target-speed prediction, semantic / BEV-semantic / depth / bounding-box auxiliary heads.  Anything else
raises like the reference does for unknown backbones (ValueError, model.py:44-46) or NotImplementedError.
"""
import math
import os
from collections import deque

import numpy as np
import torch
from torch import nn

from . import regnet
from .center_net import LidarCenterNetHead
from .transfuser import TransfuserBackbone
from .. import engine


class PerspectiveDecoder(regnet._NoForward):  # pylint: disable=protected-access
  """Parameter container of transfuser_utils.py:668-695."""

  def __init__(self, in_channels, out_channels, inter_channel_0, inter_channel_1, inter_channel_2, scale_factor_0,
               scale_factor_1):
    super().__init__()
    self.scale_factor_0 = scale_factor_0
    self.scale_factor_1 = scale_factor_1
    self.deconv1 = nn.Sequential(nn.Conv2d(in_channels, inter_channel_0, 3, 1, 1), nn.ReLU(True),
                                 nn.Conv2d(inter_channel_0, inter_channel_1, 3, 1, 1), nn.ReLU(True))
    self.deconv2 = nn.Sequential(nn.Conv2d(inter_channel_1, inter_channel_2, 3, 1, 1), nn.ReLU(True),
                                 nn.Conv2d(inter_channel_2, inter_channel_2, 3, 1, 1), nn.ReLU(True))
    self.deconv3 = nn.Sequential(nn.Conv2d(inter_channel_2, inter_channel_2, 3, 1, 1), nn.ReLU(True),
                                 nn.Conv2d(inter_channel_2, out_channels, 3, 1, 1))


class PIDController:
  """transfuser_utils.py:316-338 (host-side control, unchanged semantics)."""

  def __init__(self, k_p=1.0, k_i=0.0, k_d=0.0, n=20):
    self.k_p, self.k_i, self.k_d = k_p, k_i, k_d
    self.window = deque([0 for _ in range(n)], maxlen=n)

  def step(self, error):
    self.window.append(error)
    if len(self.window) >= 2:
      integral = np.mean(self.window)
      derivative = self.window[-1] - self.window[-2]
    else:
      integral = 0.0
      derivative = 0.0
    return self.k_p * error + self.k_i * integral + self.k_d * derivative


def valid_bev_pixels(config):
  """Camera-frustum mask of the BEV grid (transfuser_utils.py:596-665 + model.py:93-97), computed once on the host
  at construction like the reference does.  Returns (1,1,H,W) f32."""
  mpp = 1.0 / config.pixels_per_meter
  widths = torch.arange(config.min_x, config.max_x, mpp) + (mpp * 0.5)
  depths = torch.arange(config.min_y, config.max_y, mpp) + (mpp * 0.5)
  mpph = mpp * config.bev_grid_height_downsample_factor
  heights = torch.arange(config.min_z_projection, config.max_z_projection, mpph) + (mpph * 0.5)
  depths, widths, heights = torch.meshgrid(depths, widths, heights, indexing='ij')
  cloud = torch.stack((depths, widths, heights), dim=0)
  _, d, w, h = cloud.shape
  assert config.camera_rot_0[0] == config.camera_rot_0[1] == config.camera_rot_0[2] == 0.0
  t = torch.tensor(config.camera_pos).unsqueeze(1)
  c2 = cloud.view(3, -1) - t
  c2 = torch.stack((c2[1], c2[2], c2[0]))
  f = config.camera_width / (2.0 * np.tan(config.camera_fov * np.pi / 360.0))
  k = torch.from_numpy(np.array([[f, 0.0, config.camera_width / 2.0], [0.0, f, config.camera_height / 2.0],
                                 [0.0, 0.0, 1.0]])).to(dtype=torch.float32)
  c2 = k @ c2
  z = c2[2:3]
  uv = (c2[:2] / z).view(2, d, w, h)
  z = z.view(1, d, w, h)
  ok = (uv[0:1] >= 0.0) & (uv[0:1] < config.camera_width) & (uv[1:2] >= 0.0) & (uv[1:2] < config.camera_height) & (z > 0.0)
  vb = torch.max(ok.float(), dim=3)[0].unsqueeze(1)
  return torch.transpose(vb, 2, 3).contiguous()


class GRUWaypointsPredictorInterFuser(regnet._NoForward):  # pylint: disable=protected-access
  """Parameter container of model.py:839-855; forward = tfpp_planner_head."""

  def __init__(self, input_dim, waypoints, hidden_size, target_point_size):
    super().__init__()
    self.gru = torch.nn.GRU(input_size=input_dim, hidden_size=hidden_size, batch_first=True)
    if target_point_size > 0:
      self.encoder = nn.Linear(target_point_size, hidden_size)
    self.target_point_size = target_point_size
    self.hidden_size = hidden_size
    self.decoder = nn.Linear(hidden_size, 2)
    self.waypoints = waypoints


class GRUWaypointsPredictorTransFuser(regnet._NoForward):  # pylint: disable=protected-access
  """Parameter container of model.py:870-884 (the waypoint GRU of the original TransFuser: autoregressive GRUCell,
  hidden state initialised with the scene feature); forward = tfpp_gru_cell_head."""

  def __init__(self, config, pred_len, hidden_size, target_point_size):
    super().__init__()
    self.wp_decoder = nn.GRUCell(input_size=2 + target_point_size, hidden_size=hidden_size)
    self.output = nn.Linear(hidden_size, 2)
    self.config = config
    self.prediction_len = pred_len
    self.hidden_size = hidden_size


class PositionEmbeddingSine(nn.Module):
  """model.py:916-953. Input-independent for a fixed map size; built on the host once and cached by the engine."""

  def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
    super().__init__()
    self.num_pos_feats = num_pos_feats
    self.temperature = temperature
    self.normalize = normalize
    if scale is not None and normalize is False:
      raise ValueError('normalize should be True if scale is passed')
    self.scale = 2 * math.pi if scale is None else scale

  def table(self, h, w):
    """(h*w, 2*num_pos_feats) f32, row = y*w + x, channels = cat(pos_y, pos_x) (model.py:934-953)."""
    ones = torch.ones((1, h, w))
    y_embed = ones.cumsum(1, dtype=torch.float32)
    x_embed = ones.cumsum(2, dtype=torch.float32)
    if self.normalize:
      eps = 1e-6
      y_embed = y_embed / (y_embed[:, -1:, :] + eps) * self.scale
      x_embed = x_embed / (x_embed[:, :, -1:] + eps) * self.scale
    dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32)
    dim_t = self.temperature**(2 * (torch.div(dim_t, 2, rounding_mode='floor')) / self.num_pos_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).reshape(h * w, 2 * self.num_pos_feats)

  def forward(self, tensor):
    _, _, h, w = tensor.shape
    return self.table(h, w).view(1, h, w, -1).permute(0, 3, 1, 2).to(tensor.device).expand(tensor.shape[0], -1, -1, -1)


class LidarCenterNet(nn.Module):
  """The main model class (model.py:24-445) on B200 kernels."""

  def __init__(self, config):
    super().__init__()
    self.config = config
    self.speed_histogram = []
    self.make_histogram = int(os.environ.get('HISTOGRAM', 0))
    if config.backbone == 'transFuser':
      self.backbone = TransfuserBackbone(config)
    elif config.backbone == 'bev_encoder':  # model.py:42-43
      from .bev_encoder import BevEncoder  # pylint: disable=import-outside-toplevel
      self.backbone = BevEncoder(config)
    elif config.backbone == 'aim':
      raise NotImplementedError('backbone aim (camera only) is not built (SURVEY.md §8 f3)')
    else:
      raise ValueError('The chosen vision backbone does not exist. The options are: transFuser, aim, bev_encoder')
    if not (config.use_controller_input_prediction or config.use_wp_gru) or config.tp_attention or config.multi_wp_output:
      raise NotImplementedError('built planners: checkpoint + target speed (default) and / or the waypoint GRU '
                                '(use_wp_gru) through the transformer decoder; tp_attention / multi_wp_output are not')
    target_point_size = 2 if config.use_tp else 0
    self.extra_sensors = config.use_velocity or config.use_discrete_command
    if not (config.use_velocity and config.use_discrete_command and config.use_tp):
      raise NotImplementedError('default extra sensors (velocity + command + target point) only')
    if config.detect_boxes:
      self.head = LidarCenterNetHead(config)
    if config.use_semantic:
      self.semantic_decoder = PerspectiveDecoder(
          in_channels=self.backbone.num_image_features, out_channels=config.num_semantic_classes,
          inter_channel_0=config.deconv_channel_num_0, inter_channel_1=config.deconv_channel_num_1,
          inter_channel_2=config.deconv_channel_num_2,
          scale_factor_0=self.backbone.perspective_upsample_factor // config.deconv_scale_factor_0,
          scale_factor_1=self.backbone.perspective_upsample_factor // config.deconv_scale_factor_1)
    if config.use_bev_semantic:
      self.bev_semantic_decoder = nn.Sequential(
          nn.Conv2d(config.bev_features_chanels, config.bev_features_chanels, kernel_size=(3, 3), stride=1,
                    padding=(1, 1), bias=True), nn.ReLU(inplace=True),
          nn.Conv2d(config.bev_features_chanels, config.num_bev_semantic_classes, kernel_size=(1, 1), stride=1,
                    padding=0, bias=True),
          nn.Upsample(size=(config.lidar_resolution_height, config.lidar_resolution_width), mode='bilinear',
                      align_corners=False))
      vb = valid_bev_pixels(config)
      self.valid_bev_pixels = nn.Parameter(vb, requires_grad=False)
      self.valid_bev_pixels_inv = nn.Parameter(1.0 - vb, requires_grad=False)
    if config.use_depth:
      self.depth_decoder = PerspectiveDecoder(
          in_channels=self.backbone.num_image_features, out_channels=1, inter_channel_0=config.deconv_channel_num_0,
          inter_channel_1=config.deconv_channel_num_1, inter_channel_2=config.deconv_channel_num_2,
          scale_factor_0=self.backbone.perspective_upsample_factor // config.deconv_scale_factor_0,
          scale_factor_1=self.backbone.perspective_upsample_factor // config.deconv_scale_factor_1)
    d = config.gru_input_size
    if not config.transformer_decoder_join:  # model.py:184-209: the original TransFuser planner
      self._init_mlp_join(config, target_point_size)
      return
    if config.use_controller_input_prediction:
      self.target_speed_network = nn.Sequential(nn.Linear(d, d), nn.ReLU(inplace=True),
                                                nn.Linear(d, len(config.target_speeds)))
    decoder_norm = nn.LayerNorm(d)
    # nn.GELU() module + deepcopy inside nn.TransformerDecoder => the clones run F.relu (see oracle decoder_layer());
    # the engine follows the behaviour of the container it is given.
    decoder_layer = nn.TransformerDecoderLayer(d, config.num_decoder_heads, activation=nn.GELU(), batch_first=True)
    self.join = torch.nn.TransformerDecoder(decoder_layer, num_layers=config.num_transformer_decoder_layers,
                                            norm=decoder_norm)
    self.encoder_pos_encoding = PositionEmbeddingSine(d // 2, normalize=True)
    self.extra_sensor_pos_embed = nn.Parameter(torch.zeros(1, d))
    self.change_channel = nn.Conv2d(self.backbone.num_features, d, kernel_size=1)
    if config.use_wp_gru:  # model.py:165-171
      n_wp = config.pred_len // config.wp_dilation
      self.wp_query = nn.Parameter(torch.zeros(1, n_wp, d))
      self.wp_decoder = GRUWaypointsPredictorInterFuser(input_dim=d, hidden_size=config.gru_hidden_size, waypoints=n_wp,
                                                        target_point_size=target_point_size)
    if config.use_controller_input_prediction:  # model.py:173-180
      self.checkpoint_query = nn.Parameter(torch.zeros(1, config.predict_checkpoint_len + 1, d))
      self.checkpoint_decoder = GRUWaypointsPredictorInterFuser(input_dim=d, hidden_size=config.gru_hidden_size,
                                                                waypoints=config.predict_checkpoint_len,
                                                                target_point_size=target_point_size)
    self.reset_parameters()
    self.velocity_normalization = nn.BatchNorm1d(1, affine=False)
    self.extra_sensor_encoder = nn.Sequential(nn.Linear(7, 128), nn.ReLU(inplace=True), nn.Linear(128, d),
                                              nn.ReLU(inplace=True))
    self._init_common(config)

  def _init_mlp_join(self, config, target_point_size):
    """transformer_decoder_join = False (model.py:113-118,184-209,211-221): global-pooled features ++ extra-sensor
    embedding -> 3-layer MLP -> GRUWaypointsPredictorTransFuser heads + target-speed MLP on the first hidden_size
    features.  Same registration order as the reference (state_dict key order)."""
    hs = config.gru_hidden_size
    if config.use_controller_input_prediction:
      self.target_speed_network = nn.Sequential(nn.Linear(hs, hs), nn.ReLU(inplace=True),
                                                nn.Linear(hs, len(config.target_speeds)))
    join_out = hs + 2 if config.learn_origin else hs
    self.join = nn.Sequential(nn.Linear(self.backbone.num_features + config.extra_sensor_channels, 256),
                              nn.ReLU(inplace=True), nn.Linear(256, 128), nn.ReLU(inplace=True),
                              nn.Linear(128, join_out), nn.ReLU(inplace=True))
    if config.use_wp_gru:
      self.wp_decoder = GRUWaypointsPredictorTransFuser(config, pred_len=config.pred_len // config.wp_dilation,
                                                        hidden_size=hs, target_point_size=target_point_size)
    if config.use_controller_input_prediction:
      self.checkpoint_decoder = GRUWaypointsPredictorTransFuser(config, pred_len=config.predict_checkpoint_len,
                                                                hidden_size=hs, target_point_size=target_point_size)
    self.velocity_normalization = nn.BatchNorm1d(1, affine=False)
    self.extra_sensor_encoder = nn.Sequential(nn.Linear(7, 128), nn.ReLU(inplace=True),
                                              nn.Linear(128, config.extra_sensor_channels), nn.ReLU(inplace=True))
    self._init_common(config)

  def _init_common(self, config):
    self.turn_controller = PIDController(config.turn_kp, config.turn_ki, config.turn_kd, config.turn_n)
    self.speed_controller = PIDController(config.speed_kp, config.speed_ki, config.speed_kd, config.speed_n)
    self.turn_controller_direct = PIDController(config.turn_kp, config.turn_ki, config.turn_kd, config.turn_n)
    self.speed_controller_direct = PIDController(config.speed_kp, config.speed_ki, config.speed_kd, config.speed_n)
    if config.use_speed_weights:
      self.speed_weights = torch.tensor(config.target_speed_weights)
    else:
      self.speed_weights = torch.ones_like(torch.tensor(config.target_speed_weights))
    self.semantic_weights = torch.tensor(config.semantic_weights)
    self.bev_semantic_weights = torch.tensor(config.bev_semantic_weights)
    label_smoothing = config.label_smoothing_alpha if config.use_label_smoothing else 0.0
    if config.use_focal_loss or label_smoothing != 0.0:
      raise NotImplementedError('focal loss / label smoothing are off by default (config.py:211,266) and not built')
    # same buffers the reference registers through its nn.CrossEntropyLoss members (keys loss_*.weight)
    self.loss_speed = nn.CrossEntropyLoss(weight=self.speed_weights, label_smoothing=label_smoothing)
    self.loss_semantic = nn.CrossEntropyLoss(weight=self.semantic_weights, label_smoothing=label_smoothing)
    self.loss_bev_semantic = nn.CrossEntropyLoss(weight=self.bev_semantic_weights, label_smoothing=label_smoothing,
                                                 ignore_index=-1)
    self._engine = None
    self._boundary = None

  def reset_parameters(self):
    if self.config.use_wp_gru:
      nn.init.uniform_(self.wp_query)
    if self.config.use_controller_input_prediction:
      nn.init.uniform_(self.checkpoint_query)
    nn.init.uniform_(self.extra_sensor_pos_embed)

  @property
  def engine(self):
    if self._engine is None:
      object.__setattr__(self, '_engine', engine.Engine(self))
    return self._engine

  def forward(self, rgb, lidar_bev, target_point, ego_vel, command):
    """model.py:279-392: returns (pred_wp, pred_target_speed, pred_checkpoint, pred_semantic, pred_bev_semantic,
    pred_depth, pred_bounding_box, attention_weights, pred_wp_1, selected_path)."""
    if self.training and torch.is_grad_enabled():
      # the reference's train loop (train.py:776-820,898): outputs carry a grad_fn, loss.backward() reaches the parameters
      return self.boundary.forward(rgb, lidar_bev, target_point, ego_vel, command)
    return self.engine.forward(rgb, lidar_bev, target_point, ego_vel, command, training=self.training)

  @property
  def boundary(self):
    """carla_garage_b200.boundary.TrainBoundary of this model (autograd-compatible training path), built on first use."""
    if self._boundary is None:
      if getattr(self, '_trainer_owned', False):
        raise RuntimeError('this model is driven by carla_garage_b200.training.Trainer (fused step); use Trainer.step() '
                           'or build a fresh model for the autograd path')
      from ..boundary import TrainBoundary  # pylint: disable=import-outside-toplevel
      object.__setattr__(self, '_boundary', TrainBoundary(self))
    return self._boundary

  def compute_loss(self, pred_wp, pred_target_speed, pred_checkpoint, pred_semantic, pred_bev_semantic, pred_depth,
                   pred_bounding_box, pred_wp_1, selected_path, waypoint_label, target_speed_label, checkpoint_label,
                   semantic_label, bev_semantic_label, depth_label, center_heatmap_label, wh_label, yaw_class_label,
                   yaw_res_label, offset_label, velocity_label, brake_target_label, pixel_weight_label,
                   avg_factor_label):
    """model.py:394-445 (+ center_net.py:77-123) on fused loss kernels."""
    del pred_wp_1, selected_path, velocity_label, brake_target_label
    from ..boundary import compute_loss  # pylint: disable=import-outside-toplevel
    return compute_loss(self, pred_wp, waypoint_label, pred_target_speed, pred_checkpoint, pred_semantic, pred_bev_semantic, pred_depth,
                        pred_bounding_box, target_speed_label, checkpoint_label, semantic_label, bev_semantic_label,
                        depth_label, center_heatmap_label, wh_label, yaw_class_label, yaw_res_label, offset_label,
                        pixel_weight_label, avg_factor_label)

  def convert_features_to_bb_metric(self, bb_predictions):
    """model.py:447-459: decode on the GPU, threshold + image->vehicle frame on the host like the reference."""
    bboxes = self.head.get_bboxes(*bb_predictions[:5])[0]
    bboxes = bboxes[bboxes[:, -1] > self.config.bb_confidence_threshold]
    out = []
    for bbox in bboxes.detach().cpu().numpy():
      # transfuser_utils.bb_image_to_vehicle_system (transfuser_utils.py:388-406)
      bbox = bbox.copy()
      bbox[4] = -bbox[4]
      translation = np.array([-(self.config.min_x * self.config.pixels_per_meter),
                              -(self.config.min_y * self.config.pixels_per_meter)])
      bbox[:2] = bbox[:2] - translation
      bbox[0], bbox[1] = bbox[1], bbox[0]
      bbox[2], bbox[3] = bbox[3], bbox[2]
      bbox[:4] = bbox[:4] / self.config.pixels_per_meter
      out.append(bbox)
    return out

  def control_pid_direct(self, pred_target_speed, pred_angle, speed):
    """model.py:461-501."""
    if self.make_histogram:
      self.speed_histogram.append(pred_target_speed * 3.6)
    speed = speed[0].data.cpu().numpy()
    brake = pred_target_speed < 0.01
    if speed < 0.01:
      pred_angle = 0.0
    steer = round(float(np.clip(self.turn_controller_direct.step(pred_angle), -1.0, 1.0)), 3)
    if not brake and (speed / pred_target_speed) > self.config.brake_ratio:
      brake = True
    target_speed = 0.0 if brake else pred_target_speed
    delta = np.clip(target_speed - speed, 0.0, self.config.clip_delta)
    throttle = np.clip(self.speed_controller_direct.step(delta), 0.0, self.config.clip_throttle)
    if brake:
      throttle = 0.0
    return steer, throttle, brake

  def control_pid(self, waypoints, velocity):
    """model.py:503-554."""
    assert waypoints.size(0) == 1
    waypoints = waypoints[0].data.cpu().numpy()
    speed = velocity[0].data.cpu().numpy()
    one_second = int(self.config.carla_fps // (self.config.wp_dilation * self.config.data_save_freq))
    half_second = one_second // 2
    desired_speed = np.linalg.norm(waypoints[half_second - 1] - waypoints[one_second - 1]) * 2.0
    if self.make_histogram:
      self.speed_histogram.append(desired_speed * 3.6)
    brake = (desired_speed < self.config.brake_speed) or ((speed / desired_speed) > self.config.brake_ratio)
    delta = np.clip(desired_speed - speed, 0.0, self.config.clip_delta)
    throttle = np.clip(self.speed_controller.step(delta), 0.0, self.config.clip_throttle)
    throttle = throttle if not brake else 0.0
    aim_distance = self.config.aim_distance_slow if desired_speed < self.config.aim_distance_threshold else \
        self.config.aim_distance_fast
    aim_index = waypoints.shape[0] - 1
    for index, wp in enumerate(waypoints):
      if np.linalg.norm(wp) >= aim_distance:
        aim_index = index
        break
    aim = waypoints[aim_index]
    angle = np.degrees(np.arctan2(aim[1], aim[0])) / 90.0
    if speed < 0.01 or brake:
      angle = 0.0
    steer = np.clip(self.turn_controller.step(angle), -1.0, 1.0)
    return steer, throttle, brake

  def create_optimizer_groups(self, weight_decay):
    """model.py:556-645: same name/type rules, so the decay / no-decay split is identical."""
    decay, no_decay = set(), set()
    whitelist = (torch.nn.Linear, torch.nn.Conv2d)
    blacklist = (torch.nn.LayerNorm, torch.nn.Embedding, torch.nn.BatchNorm2d)
    for mn, m in self.named_modules():
      for pn, _ in m.named_parameters():
        fpn = f'{mn}.{pn}' if mn else pn
        if pn.endswith('bias'):
          no_decay.add(fpn)
        elif pn.endswith('weight') and isinstance(m, whitelist):
          decay.add(fpn)
        elif pn.endswith('weight') and isinstance(m, blacklist):
          no_decay.add(fpn)
        elif pn.endswith('weight') and 'conv.' in pn:
          decay.add(fpn)
        elif pn.endswith('weight') and ('.bn' in pn or '.ln' in pn):
          no_decay.add(fpn)
        elif pn.endswith('weight') and 'downsample.0.weight' in pn:
          decay.add(fpn)
        elif pn.endswith('weight') and 'downsample.1.weight' in pn:
          no_decay.add(fpn)
        elif pn.endswith('weight') and ('.attn' in pn or 'channel_to_' in pn or '.mlp' in pn or
                                        'target_speed_network' in pn):
          decay.add(fpn)
        elif pn.endswith('weight') and 'join.' in pn and '.norm' not in pn:
          decay.add(fpn)
        elif pn.endswith('weight') and 'join.' in pn and '.norm' in pn:
          no_decay.add(fpn)
        elif pn.endswith('_ih') or pn.endswith('_hh'):
          no_decay.add(fpn)
        elif pn.endswith('_emb') or '_token' in pn or pn.endswith('_embed'):
          no_decay.add(fpn)
        elif 'bias_ih_l0' in pn or 'bias_hh_l0' in pn:
          no_decay.add(fpn)
        elif 'weight_ih_l0' in pn or 'weight_hh_l0' in pn:
          decay.add(fpn)
        elif '_query' in pn or 'weight_hh_l0' in pn:
          no_decay.add(fpn)
        elif 'valid_bev_pixels' in pn:
          no_decay.add(fpn)
    param_dict = dict(self.named_parameters())
    inter, union = decay & no_decay, decay | no_decay
    assert len(inter) == 0, f'parameters {inter} made it into both decay/no_decay sets!'
    assert len(param_dict.keys() - union) == 0, f'parameters {param_dict.keys() - union} were not separated'
    return [{'params': [param_dict[pn] for pn in sorted(decay)], 'weight_decay': weight_decay},
            {'params': [param_dict[pn] for pn in sorted(no_decay)], 'weight_decay': 0.0}]

  def init_visualization(self):
    """model.py:647-663 needs the CARLA map renderer; debug visualisation is outside the hot path."""
    if self.config.debug:
      raise NotImplementedError('debug visualisation needs a CARLA server (out of scope, SURVEY.md §2a)')

  def visualize_model(self, *args, **kwargs):
    raise NotImplementedError('debug visualisation (model.py:665-836) is outside the hot path')
