"""TransfuserBackbone / GPT with the reference's constructor, attributes and state_dict keys
(team_code/transfuser.py:16-137,260-299,342-402); the forward pass runs on the sm_100a kernels via
carla_garage_b200.engine.  Only the default TransFuser++ configuration is built here (regnety_032 branches, 2-D LiDAR
BEV); other ``lidar_architecture`` values are outside this path."""
import torch
from torch import nn

from . import regnet
from .. import engine


class SelfAttention(regnet._NoForward):  # pylint: disable=protected-access
  """Parameter container of transfuser.py:342-360."""

  def __init__(self, n_embd, n_head, attn_pdrop, resid_pdrop):
    super().__init__()
    assert n_embd % n_head == 0
    self.key = nn.Linear(n_embd, n_embd)
    self.query = nn.Linear(n_embd, n_embd)
    self.value = nn.Linear(n_embd, n_embd)
    self.attn_drop = nn.Dropout(attn_pdrop)
    self.resid_drop = nn.Dropout(resid_pdrop)
    self.proj = nn.Linear(n_embd, n_embd)
    self.n_head = n_head


class Block(regnet._NoForward):  # pylint: disable=protected-access
  """transfuser.py:383-396."""

  def __init__(self, n_embd, n_head, block_exp, attn_pdrop, resid_pdrop):
    super().__init__()
    self.ln1 = nn.LayerNorm(n_embd)
    self.ln2 = nn.LayerNorm(n_embd)
    self.attn = SelfAttention(n_embd, n_head, attn_pdrop, resid_pdrop)
    self.mlp = nn.Sequential(
        nn.Linear(n_embd, block_exp * n_embd),
        nn.ReLU(True),
        nn.Linear(block_exp * n_embd, n_embd),
        nn.Dropout(resid_pdrop),
    )


class GPT(regnet._NoForward):  # pylint: disable=protected-access
  """transfuser.py:260-299 (parameters, init); forward = engine.gpt_forward."""

  def __init__(self, n_embd, config, lidar_video=False, lidar_time_frames=1):
    super().__init__()
    if lidar_video:
      raise NotImplementedError('temporal LiDAR backbones are outside the TransFuser++ hot path (SURVEY.md §8f)')
    self.n_embd = n_embd
    self.seq_len = 1
    self.config = config
    self.lidar_time_frames = lidar_time_frames
    self.pos_emb = nn.Parameter(
        torch.zeros(1, config.img_vert_anchors * config.img_horz_anchors +
                    lidar_time_frames * config.lidar_vert_anchors * config.lidar_horz_anchors, n_embd))
    self.drop = nn.Dropout(config.embd_pdrop)
    self.blocks = nn.Sequential(*[
        Block(n_embd, config.n_head, config.block_exp, config.attn_pdrop, config.resid_pdrop)
        for _ in range(config.n_layer)
    ])
    self.ln_f = nn.LayerNorm(n_embd)
    self.apply(self._init_weights)

  def _init_weights(self, module):
    if isinstance(module, nn.Linear):
      module.weight.data.normal_(mean=self.config.gpt_linear_layer_init_mean, std=self.config.gpt_linear_layer_init_std)
      if module.bias is not None:
        module.bias.data.zero_()
    elif isinstance(module, nn.LayerNorm):
      module.bias.data.zero_()
      module.weight.data.fill_(self.config.gpt_layer_norm_init_weight)


class TransfuserBackbone(nn.Module):
  """Multi-scale fusion transformer for image + LiDAR features (transfuser.py:16-257), CUDA-native forward."""

  def __init__(self, config):
    super().__init__()
    self.config = config
    if config.image_architecture != 'regnety_032' or config.lidar_architecture != 'regnety_032':
      raise NotImplementedError('carla_garage_b200 builds the default TransFuser++ branches (regnety_032) only')
    self.image_encoder = regnet.RegNetY032Features(in_chans=3)
    self.lidar_video = False
    in_channels = (2 if config.use_ground_plane else 1) * config.lidar_seq_len
    self.lidar_encoder = regnet.RegNetY032Features(in_chans=in_channels)
    self.avgpool_img = nn.AdaptiveAvgPool2d((config.img_vert_anchors, config.img_horz_anchors))
    self.avgpool_lidar = nn.AdaptiveAvgPool2d((config.lidar_vert_anchors, config.lidar_horz_anchors))
    self.global_pool_lidar = nn.AdaptiveAvgPool2d(output_size=1)
    self.global_pool_img = nn.AdaptiveAvgPool2d(output_size=1)
    start_index = 1  # RegNet has a stem return layer (transfuser.py:61-64)
    info_i = self.image_encoder.feature_info.info
    info_l = self.lidar_encoder.feature_info.info
    self.transformers = nn.ModuleList([
        GPT(n_embd=info_i[start_index + i]['num_chs'], config=config, lidar_video=False, lidar_time_frames=1)
        for i in range(4)
    ])
    self.lidar_channel_to_img = nn.ModuleList(
        [nn.Conv2d(info_l[start_index + i]['num_chs'], info_i[start_index + i]['num_chs'], kernel_size=1)
         for i in range(4)])
    self.img_channel_to_lidar = nn.ModuleList(
        [nn.Conv2d(info_i[start_index + i]['num_chs'], info_l[start_index + i]['num_chs'], kernel_size=1)
         for i in range(4)])
    self.num_image_features = info_i[start_index + 3]['num_chs']
    self.perspective_upsample_factor = info_i[start_index + 3]['reduction'] // config.perspective_downsample_factor
    if config.transformer_decoder_join:
      self.num_features = info_l[start_index + 3]['num_chs']
    else:  # transfuser.py:102-111: globally pooled image + LiDAR features feed an MLP join
      if not config.add_features:
        raise NotImplementedError('add_features=False (concatenated global features) is not built')
      self.lidar_to_img_features_end = nn.Linear(info_l[start_index + 3]['num_chs'], info_i[start_index + 3]['num_chs'])
      self.num_features = info_i[start_index + 3]['num_chs']
    channel = config.bev_features_chanels
    self.relu = nn.ReLU(inplace=True)
    if config.detect_boxes or config.use_bev_semantic:
      self.upsample = nn.Upsample(scale_factor=config.bev_upsample_factor, mode='bilinear', align_corners=False)
      self.upsample2 = nn.Upsample(size=(config.lidar_resolution_height // config.bev_down_sample_factor,
                                         config.lidar_resolution_width // config.bev_down_sample_factor),
                                   mode='bilinear',
                                   align_corners=False)
      self.up_conv5 = nn.Conv2d(channel, channel, (3, 3), padding=1)
      self.up_conv4 = nn.Conv2d(channel, channel, (3, 3), padding=1)
      self.c5_conv = nn.Conv2d(info_l[start_index + 3]['num_chs'], channel, (1, 1))

  def forward(self, image, lidar):
    """Same contract as transfuser.py:139-205: NCHW f32 in, (features, fused_features, image_feature_grid) NCHW f32
    out.  (LidarCenterNet consumes the NHWC bf16 internals directly and never converts.)"""
    eng = engine.Engine.for_backbone(self)
    feats, fused, grid = eng.backbone_forward(image, lidar, training=self.training)
    from .. import ops  # pylint: disable=import-outside-toplevel
    if not self.config.transformer_decoder_join:
      fused = fused.float()   # (B, num_features): globally pooled image + LiDAR features (transfuser.py:188-197)
    else:
      fused = ops.nhwc_to_nchw(fused)
    return (ops.nhwc_to_nchw(feats) if feats is not None else None, fused,
            ops.nhwc_to_nchw(grid) if grid is not None else None)
