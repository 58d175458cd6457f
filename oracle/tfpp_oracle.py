"""ORACLE — test infrastructure only.

CPU fp32 restatement (plain ``torch.nn.functional`` on CPU tensors + numpy for the integer part) of the
TransFuser++ forward / loss of autonomousvision/carla_garage, written as *functions of a state_dict* so that it
shares no code with the product modules in ``carla_garage_b200/``.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it.

Every function cites the reference lines it restates (paths relative to /root/reference/team_code).
The RegNetY arithmetic (timm 0.6.7, not in the reference tree) is restated per ``oracle/regnety.py``'s header.

Parity status: the reference has NO tests or golden vectors for this path => "parity unpinned"
by the reference itself.  The oracle is instead pinned to *outputs of the reference run in the build container*:
``tests/golden/make_golden.py`` imports the unmodified reference modules (with ``carla_garage_b200.compat`` stubs),
runs them on seeded inputs/weights and commits the vectors; ``tests/test_oracle.py`` checks this file against
those vectors everywhere and against the live reference when /root/reference exists.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

DEFAULT_CFG = dict(
    # config.py:131-138,481 LiDAR voxelisation
    min_x=-32, max_x=32, min_y=-32, max_y=32, pixels_per_meter=4.0, hist_max_per_pixel=5, lidar_split_height=0.2,
    max_height_lidar=100.0, use_ground_plane=False,
    # config.py:333-357 fusion transformer
    img_vert_anchors=8, img_horz_anchors=32, lidar_vert_anchors=8, lidar_horz_anchors=8, n_head=4, n_layer=2,
    block_exp=4,
    # config.py:343-347 BEV pyramid
    bev_features_chanels=64, bev_upsample_factor=2, bev_down_sample_factor=4, lidar_resolution_height=256,
    lidar_resolution_width=256,
    # config.py:321-322,366,468-469 planner
    gru_hidden_size=64, gru_input_size=256, predict_checkpoint_len=10, num_transformer_decoder_layers=6,
    num_decoder_heads=8, decoder_activation='relu',  # see decoder_layer(): deepcopy turns nn.GELU() into F.relu
    # config.py:451-458 perspective decoders: scale factors are 32 // 4 = 8 and 32 // 8 = 4 (model.py:71-72)
    perspective_scale_0=8, perspective_scale_1=4,
    # config.py:307-317 detector
    num_dir_bins=12, top_k_center_keypoints=100, center_net_max_pooling_kernel=3, num_bb_classes=4,
    # config.py:158 class weights (use_speed_weights=True, label smoothing off: config.py:262-266)
    target_speed_weights=[0.866605263873406, 7.4527377240841775, 1.2281629310898465, 0.5269622904065803],
)


# ---------------------------------------------------------------------------------------------------------------
# a1  LiDAR point cloud -> BEV histogram  (data.py:873-906)
# ---------------------------------------------------------------------------------------------------------------
def lidar_to_histogram_features(lidar, use_ground_plane=False, cfg=None):
  """data.py:873-906.  ``np.histogramdd`` restated as explicit integer binning.

  Edges are linspace(-32, 32, 257) = exact multiples of 0.25, bins are half-open [e_i, e_{i+1}) with the last bin
  closed (numpy histogram semantics), so bin(x) = floor(4x) + 128 for -32 <= x < 32 and 255 for x == 32.
  Comparisons happen in the dtype of ``lidar`` (float32 clouds compare against float32(0.2)), like numpy does.
  Output index is [channel, y_bin, x_bin] (the ``.T`` at data.py:893).
  """
  cfg = cfg or DEFAULT_CFG
  lidar = np.asarray(lidar)
  dt = lidar.dtype.type
  ppm = int(cfg['pixels_per_meter'])
  nx = (cfg['max_x'] - cfg['min_x']) * ppm
  ny = (cfg['max_y'] - cfg['min_y']) * ppm

  def splat(pc):
    x, y = pc[:, 0].astype(np.float64), pc[:, 1].astype(np.float64)
    ok = (x >= cfg['min_x']) & (x <= cfg['max_x']) & (y >= cfg['min_y']) & (y <= cfg['max_y'])
    bx = np.minimum(np.floor(x[ok] * ppm).astype(np.int64) - cfg['min_x'] * ppm, nx - 1)
    by = np.minimum(np.floor(y[ok] * ppm).astype(np.int64) - cfg['min_y'] * ppm, ny - 1)
    hist = np.zeros((nx, ny), dtype=np.int64)
    np.add.at(hist, (bx, by), 1)
    hist = np.minimum(hist, cfg['hist_max_per_pixel'])
    return (hist / cfg['hist_max_per_pixel']).T

  lidar = lidar[lidar[:, 2] < dt(cfg['max_height_lidar'])]
  below = lidar[lidar[:, 2] <= dt(cfg['lidar_split_height'])]
  above = lidar[lidar[:, 2] > dt(cfg['lidar_split_height'])]
  feats = [splat(below), splat(above)] if use_ground_plane else [splat(above)]
  return np.stack(feats, axis=0).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# a2/a3  image normalisation + RegNetY-3.2GF  (transfuser_utils.py:542-551; timm 0.6.7 regnet.py)
# ---------------------------------------------------------------------------------------------------------------
def normalize_imagenet(x):
  """transfuser_utils.py:542-551."""
  mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
  std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
  return ((x / 255.0) - mean) / std


def _bn(sd, p, x, training):
  # timm BatchNormAct2d == nn.BatchNorm2d(eps=1e-5, momentum=0.1); training uses batch stats (running stats are
  # not updated here: the oracle is stateless, SURVEY.md §8a' BatchNorm2d row).
  return F.batch_norm(x, sd[p + '.running_mean'].clone(), sd[p + '.running_var'].clone(), sd[p + '.weight'],
                      sd[p + '.bias'], training=training, momentum=0.1, eps=1e-5)


def _conv_bn(sd, p, x, training, stride=1, groups=1, act=True):
  w = sd[p + '.conv.weight']
  x = F.conv2d(x, w, None, stride=stride, padding=w.shape[-1] // 2, groups=groups)
  x = _bn(sd, p + '.bn', x, training)
  return F.relu(x) if act else x


def regnet_block(sd, p, x, training, stride, group_width=24):
  """timm 0.6.7 regnet.py Bottleneck.forward."""
  shortcut = x
  if (p + '.downsample.conv.weight') in sd:
    shortcut = _conv_bn(sd, p + '.downsample', x, training, stride=stride, act=False)
  y = _conv_bn(sd, p + '.conv1', x, training)
  y = _conv_bn(sd, p + '.conv2', y, training, stride=stride, groups=y.shape[1] // group_width)
  s = y.mean((2, 3), keepdim=True)
  s = F.conv2d(F.relu(F.conv2d(s, sd[p + '.se.fc1.weight'], sd[p + '.se.fc1.bias'])), sd[p + '.se.fc2.weight'],
               sd[p + '.se.fc2.bias'])
  y = y * torch.sigmoid(s)
  y = _conv_bn(sd, p + '.conv3', y, training, act=False)
  return F.relu(y + shortcut)


def regnet_stage(sd, p, x, training):
  j = 1
  while (p + f'.b{j}.conv1.conv.weight') in sd:
    x = regnet_block(sd, p + f'.b{j}', x, training, stride=2 if j == 1 else 1)
    j += 1
  return x


# ---------------------------------------------------------------------------------------------------------------
# a5  GPT fusion transformer  (transfuser.py:301-402)
# ---------------------------------------------------------------------------------------------------------------
def _drop(dropout, t, pdrop):
  """A dropout site.  ``dropout`` is None (eval / parity runs without dropout) or a callable (tensor, p) -> tensor that
  applies the site's mask, e.g. oracle.philox.DropoutStream, which reproduces the engine's counter-based masks."""
  return t if dropout is None else dropout(t, pdrop)


def self_attention(sd, p, x, n_head, dropout=None, attn_pdrop=0.1, resid_pdrop=0.1):
  """transfuser.py:362-380; attn_drop on the probabilities (:374), resid_drop on the projection (:379)."""
  b, t, c = x.shape
  hd = c // n_head
  k = F.linear(x, sd[p + '.key.weight'], sd[p + '.key.bias']).view(b, t, n_head, hd).transpose(1, 2)
  q = F.linear(x, sd[p + '.query.weight'], sd[p + '.query.bias']).view(b, t, n_head, hd).transpose(1, 2)
  v = F.linear(x, sd[p + '.value.weight'], sd[p + '.value.bias']).view(b, t, n_head, hd).transpose(1, 2)
  att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(hd))
  att = F.softmax(att, dim=-1)
  att = _drop(dropout, att, attn_pdrop)
  y = (att @ v).transpose(1, 2).contiguous().view(b, t, c)
  return _drop(dropout, F.linear(y, sd[p + '.proj.weight'], sd[p + '.proj.bias']), resid_pdrop)


def gpt_block(sd, p, x, n_head, dropout=None, attn_pdrop=0.1, resid_pdrop=0.1):
  """transfuser.py:383-402 (nn.Dropout(resid_pdrop) closes the MLP, :395)."""
  c = x.shape[-1]
  h = F.layer_norm(x, (c,), sd[p + '.ln1.weight'], sd[p + '.ln1.bias'], 1e-5)
  x = x + self_attention(sd, p + '.attn', h, n_head, dropout, attn_pdrop, resid_pdrop)
  h = F.layer_norm(x, (c,), sd[p + '.ln2.weight'], sd[p + '.ln2.bias'], 1e-5)
  h = F.relu(F.linear(h, sd[p + '.mlp.0.weight'], sd[p + '.mlp.0.bias']))
  return x + _drop(dropout, F.linear(h, sd[p + '.mlp.2.weight'], sd[p + '.mlp.2.bias']), resid_pdrop)


def gpt(sd, p, image_tensor, lidar_tensor, cfg, dropout=None):
  """transfuser.py:301-339, non-video branch; self.drop on pos_emb + tokens (:325)."""
  bz, c, img_h, img_w = image_tensor.shape
  lidar_h, lidar_w = lidar_tensor.shape[2:4]
  it = image_tensor.permute(0, 2, 3, 1).contiguous().view(bz, -1, c)
  lt = lidar_tensor.permute(0, 2, 3, 1).contiguous().view(bz, -1, c)
  x = _drop(dropout, sd[p + '.pos_emb'] + torch.cat((it, lt), dim=1), cfg.get('embd_pdrop', 0.1))
  for l in range(cfg['n_layer']):
    x = gpt_block(sd, p + f'.blocks.{l}', x, cfg['n_head'], dropout, cfg.get('attn_pdrop', 0.1),
                  cfg.get('resid_pdrop', 0.1))
  x = F.layer_norm(x, (c,), sd[p + '.ln_f.weight'], sd[p + '.ln_f.bias'], 1e-5)
  n_img = img_h * img_w
  img_out = x[:, :n_img].view(bz, img_h, img_w, -1).permute(0, 3, 1, 2).contiguous()
  lid_out = x[:, n_img:].view(bz, lidar_h, lidar_w, -1).permute(0, 3, 1, 2).contiguous()
  return img_out, lid_out


def fuse_features(sd, p, image_features, lidar_features, i, cfg, dropout=None):
  """transfuser.py:222-257."""
  img_e = F.adaptive_avg_pool2d(image_features, (cfg['img_vert_anchors'], cfg['img_horz_anchors']))
  lid_e = F.adaptive_avg_pool2d(lidar_features, (cfg['lidar_vert_anchors'], cfg['lidar_horz_anchors']))
  lid_e = F.conv2d(lid_e, sd[p + f'.lidar_channel_to_img.{i}.weight'], sd[p + f'.lidar_channel_to_img.{i}.bias'])
  img_l, lid_l = gpt(sd, p + f'.transformers.{i}', img_e, lid_e, cfg, dropout)
  lid_l = F.conv2d(lid_l, sd[p + f'.img_channel_to_lidar.{i}.weight'], sd[p + f'.img_channel_to_lidar.{i}.bias'])
  img_l = F.interpolate(img_l, size=image_features.shape[2:], mode='bilinear', align_corners=False)
  lid_l = F.interpolate(lid_l, size=lidar_features.shape[2:], mode='bilinear', align_corners=False)
  return image_features + img_l, lidar_features + lid_l


def backbone_forward(sd, image, lidar, cfg, training=False, p='backbone', taps=None, dropout=None):
  """TransfuserBackbone.forward, transfuser.py:139-205 (transformer_decoder_join, detect_boxes, use_semantic)."""
  x_img = normalize_imagenet(image)
  x_lid = lidar
  x_img = _conv_bn(sd, p + '.image_encoder.stem', x_img, training, stride=2)
  x_lid = _conv_bn(sd, p + '.lidar_encoder.stem', x_lid, training, stride=2)
  if taps is not None:
    taps['img_stem'], taps['lid_stem'] = x_img, x_lid
  for i in range(4):
    x_img = regnet_stage(sd, p + f'.image_encoder.s{i + 1}', x_img, training)
    x_lid = regnet_stage(sd, p + f'.lidar_encoder.s{i + 1}', x_lid, training)
    if taps is not None:
      taps[f'img_s{i + 1}_pre'], taps[f'lid_s{i + 1}_pre'] = x_img, x_lid
    x_img, x_lid = fuse_features(sd, p, x_img, x_lid, i, cfg, dropout)
    if taps is not None:
      taps[f'img_s{i + 1}'], taps[f'lid_s{i + 1}'] = x_img, x_lid
  # top_down, transfuser.py:131-137
  p5 = F.relu(F.conv2d(x_lid, sd[p + '.c5_conv.weight'], sd[p + '.c5_conv.bias']))
  p4 = F.interpolate(p5, scale_factor=cfg['bev_upsample_factor'], mode='bilinear', align_corners=False)
  p4 = F.relu(F.conv2d(p4, sd[p + '.up_conv5.weight'], sd[p + '.up_conv5.bias'], padding=1))
  size = (cfg['lidar_resolution_height'] // cfg['bev_down_sample_factor'],
          cfg['lidar_resolution_width'] // cfg['bev_down_sample_factor'])
  p3 = F.interpolate(p4, size=size, mode='bilinear', align_corners=False)
  p3 = F.relu(F.conv2d(p3, sd[p + '.up_conv4.weight'], sd[p + '.up_conv4.bias'], padding=1))
  if not cfg.get('transformer_decoder_join', True):
    # transfuser.py:188-197 (add_features): global pools, lidar_to_img_features_end, sum
    gi = torch.flatten(F.adaptive_avg_pool2d(x_img, 1), 1)
    gl = torch.flatten(F.adaptive_avg_pool2d(x_lid, 1), 1)
    gl = F.linear(gl, sd[p + '.lidar_to_img_features_end.weight'], sd[p + '.lidar_to_img_features_end.bias'])
    return p3, gi + gl, x_img
  return p3, x_lid, x_img


# ---------------------------------------------------------------------------------------------------------------
# a8-a10  planner: memory tokens, 6-layer decoder, GRU, target speed  (model.py:299-358,839-867,916-953)
# ---------------------------------------------------------------------------------------------------------------
def position_embedding_sine(bs, h, w, num_pos_feats=128, temperature=10000, scale=2 * math.pi):
  """model.py:934-953 with normalize=True."""
  not_mask = torch.ones((bs, h, w))
  y_embed = not_mask.cumsum(1, dtype=torch.float32)
  x_embed = not_mask.cumsum(2, dtype=torch.float32)
  eps = 1e-6
  y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
  x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
  dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
  dim_t = temperature**(2 * (torch.div(dim_t, 2, rounding_mode='floor')) / num_pos_feats)
  pos_x = x_embed[:, :, :, None] / dim_t
  pos_y = y_embed[:, :, :, None] / dim_t
  pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
  pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
  return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def _mha(sd, p, q_in, kv_in, n_head, dropout=None, pdrop=0.1):
  """torch.nn.MultiheadAttention (batch_first, no masks; dropout on the attention weights) from its in_proj/out_proj
  parameters."""
  d = q_in.shape[-1]
  w, b = sd[p + '.in_proj_weight'], sd[p + '.in_proj_bias']
  q = F.linear(q_in, w[:d], b[:d])
  k = F.linear(kv_in, w[d:2 * d], b[d:2 * d])
  v = F.linear(kv_in, w[2 * d:], b[2 * d:])
  bs, tq, _ = q.shape
  tk = k.shape[1]
  hd = d // n_head
  q = q.view(bs, tq, n_head, hd).transpose(1, 2)
  k = k.view(bs, tk, n_head, hd).transpose(1, 2)
  v = v.view(bs, tk, n_head, hd).transpose(1, 2)
  att = _drop(dropout, F.softmax((q @ k.transpose(-2, -1)) / math.sqrt(hd), dim=-1), pdrop)
  y = (att @ v).transpose(1, 2).reshape(bs, tq, d)
  return F.linear(y, sd[p + '.out_proj.weight'], sd[p + '.out_proj.bias'])


def decoder_layer(sd, p, x, mem, n_head, activation='relu', dropout=None, pdrop=0.1):
  """nn.TransformerDecoderLayer(d, heads, activation=nn.GELU(), batch_first, norm_first=False), model.py:137-143.

  OBSERVED BEHAVIOUR (run here, torch 2.11, and by code inspection identical in the pinned torch 1.12.1): the
  ``nn.GELU()`` *module* passed as ``activation`` lands in ``_modules``; ``nn.TransformerDecoder`` deep-copies the
  layer (``_get_clones``), ``TransformerDecoderLayer.__setstate__`` does not find 'activation' in ``__dict__`` and
  injects ``F.relu``, which then shadows the module.  Every ``join.layers[i].activation`` of the reference model
  is therefore ``F.relu`` - the feed-forward non-linearity the reference really computes is ReLU, not GELU.  The
  oracle follows the behaviour (tests/test_oracle.py::test_oracle_vs_live_reference pins it); 'gelu' is kept as an
  option for a torch that fixes the quirk."""
  d = x.shape[-1]
  act = F.relu if activation == 'relu' else F.gelu
  # torch TransformerDecoderLayer (norm_first=False): x = norm1(x + dropout1(sa)); x = norm2(x + dropout2(mha));
  # x = norm3(x + dropout3(linear2(dropout(act(linear1(x)))))); the two attentions drop their probabilities too
  sa = _drop(dropout, _mha(sd, p + '.self_attn', x, x, n_head, dropout, pdrop), pdrop)
  x = F.layer_norm(x + sa, (d,), sd[p + '.norm1.weight'], sd[p + '.norm1.bias'])
  ca = _drop(dropout, _mha(sd, p + '.multihead_attn', x, mem, n_head, dropout, pdrop), pdrop)
  x = F.layer_norm(x + ca, (d,), sd[p + '.norm2.weight'], sd[p + '.norm2.bias'])
  h = _drop(dropout, act(F.linear(x, sd[p + '.linear1.weight'], sd[p + '.linear1.bias'])), pdrop)
  h = _drop(dropout, F.linear(h, sd[p + '.linear2.weight'], sd[p + '.linear2.bias']), pdrop)
  return F.layer_norm(x + h, (d,), sd[p + '.norm3.weight'], sd[p + '.norm3.bias'])


def gru_waypoints(sd, p, x, target_point):
  """GRUWaypointsPredictorInterFuser.forward, model.py:857-867; nn.GRU gate order (r,z,n)."""
  h = F.linear(target_point, sd[p + '.encoder.weight'], sd[p + '.encoder.bias'])
  w_ih, w_hh = sd[p + '.gru.weight_ih_l0'], sd[p + '.gru.weight_hh_l0']
  b_ih, b_hh = sd[p + '.gru.bias_ih_l0'], sd[p + '.gru.bias_hh_l0']
  hs = h.shape[-1]
  outs = []
  for t in range(x.shape[1]):
    gi = F.linear(x[:, t], w_ih, b_ih)
    gh = F.linear(h, w_hh, b_hh)
    r = torch.sigmoid(gi[:, :hs] + gh[:, :hs])
    z = torch.sigmoid(gi[:, hs:2 * hs] + gh[:, hs:2 * hs])
    n = torch.tanh(gi[:, 2 * hs:] + r * gh[:, 2 * hs:])
    h = (1 - z) * n + z * h
    outs.append(h)
  out = torch.stack(outs, dim=1)
  out = F.linear(out, sd[p + '.decoder.weight'], sd[p + '.decoder.bias'])
  return torch.cumsum(out, 1)


def perspective_decoder(sd, p, x, cfg):
  """transfuser_utils.py:697-704."""

  def c(name, t, relu=True):
    t = F.conv2d(t, sd[f'{p}.{name}.weight'], sd[f'{p}.{name}.bias'], padding=1)
    return F.relu(t) if relu else t

  x = c('deconv1.2', c('deconv1.0', x))
  x = F.interpolate(x, scale_factor=cfg['perspective_scale_0'], mode='bilinear', align_corners=False)
  x = c('deconv2.2', c('deconv2.0', x))
  x = F.interpolate(x, scale_factor=cfg['perspective_scale_1'], mode='bilinear', align_corners=False)
  return c('deconv3.2', c('deconv3.0', x), relu=False)


def center_net_head(sd, p, feat):
  """center_net.py:49-75 (single-frame: no velocity / brake heads)."""

  def head(name):
    t = F.relu(F.conv2d(feat, sd[f'{p}.{name}.0.weight'], sd[f'{p}.{name}.0.bias'], padding=1))
    return F.conv2d(t, sd[f'{p}.{name}.2.weight'], sd[f'{p}.{name}.2.bias'])

  return (head('heatmap_head').sigmoid(), head('wh_head'), head('offset_head'), head('yaw_class_head'),
          head('yaw_res_head'), None, None)


def planner(sd, fused, target_point, ego_vel, command, cfg=None, training=False, taps=None, dropout=None):
  """model.py:299-358: change_channel + sine position encoding -> 64 memory tokens, extra-sensor token, 6-layer decoder
  over the 11 learned queries, GRU checkpoints + target-speed logits.  Returns (pred_checkpoint, pred_target_speed)."""
  cfg = cfg or DEFAULT_CFG
  bs = fused.shape[0]
  # model.py:301-303
  f = F.conv2d(fused, sd['change_channel.weight'], sd['change_channel.bias'])
  f = f + position_embedding_sine(bs, f.shape[2], f.shape[3], cfg['gru_input_size'] // 2)
  f = torch.flatten(f, start_dim=2)
  # model.py:308-319
  vel = F.batch_norm(ego_vel, sd['velocity_normalization.running_mean'].clone(),
                     sd['velocity_normalization.running_var'].clone(), None, None, training=training, momentum=0.1,
                     eps=1e-5)
  es = torch.cat([vel, command], dim=1)
  es = F.relu(F.linear(es, sd['extra_sensor_encoder.0.weight'], sd['extra_sensor_encoder.0.bias']))
  es = F.relu(F.linear(es, sd['extra_sensor_encoder.2.weight'], sd['extra_sensor_encoder.2.bias']))
  es = es + sd['extra_sensor_pos_embed'].repeat(bs, 1)
  mem = torch.cat((f, es.unsqueeze(2)), dim=2).permute(0, 2, 1)
  if taps is not None:
    taps['memory'] = mem
  def join(queries):  # model.py:352  nn.TransformerDecoder + final norm
    x = queries.repeat(bs, 1, 1)
    for l in range(cfg['num_transformer_decoder_layers']):
      x = decoder_layer(sd, f'join.layers.{l}', x, mem, cfg['num_decoder_heads'], cfg.get('decoder_activation', 'relu'),
                        dropout, cfg.get('decoder_pdrop', 0.1))
    return F.layer_norm(x, (x.shape[-1],), sd['join.norm.weight'], sd['join.norm.bias'])

  pred_wp = None
  if cfg.get('use_wp_gru', False):  # model.py:325-337 (single waypoint output)
    pred_wp = gru_waypoints(sd, 'wp_decoder', join(sd['wp_query']), target_point)
  x = join(sd['checkpoint_query'])
  if taps is not None:
    taps['joined'] = x
  n = cfg['predict_checkpoint_len']
  pred_checkpoint = gru_waypoints(sd, 'checkpoint_decoder', x[:, :n], target_point)
  ts = x[:, n]
  pred_target_speed = F.linear(F.relu(F.linear(ts, sd['target_speed_network.0.weight'],
                                               sd['target_speed_network.0.bias'])),
                               sd['target_speed_network.2.weight'], sd['target_speed_network.2.bias'])
  if cfg.get('use_wp_gru', False):
    return pred_checkpoint, pred_target_speed, pred_wp
  return pred_checkpoint, pred_target_speed


def gru_waypoints_transfuser(sd, p, z, target_point, steps, cfg):
  """GRUWaypointsPredictorTransFuser.forward, model.py:886-913 (learn_origin, use_tp): autoregressive nn.GRUCell."""
  hs = cfg['gru_hidden_size']
  if cfg.get('learn_origin', 1):
    x, z = z[:, hs:hs + 2], z[:, :hs]
  else:
    x = torch.zeros((z.shape[0], 2), dtype=z.dtype)
  w_ih, w_hh = sd[p + '.wp_decoder.weight_ih'], sd[p + '.wp_decoder.weight_hh']
  b_ih, b_hh = sd[p + '.wp_decoder.bias_ih'], sd[p + '.wp_decoder.bias_hh']
  out = []
  for _ in range(steps):
    x_in = torch.cat([x, target_point], dim=1)
    gi, gh = F.linear(x_in, w_ih, b_ih), F.linear(z, w_hh, b_hh)
    i_r, i_z, i_n = gi.chunk(3, 1)
    h_r, h_z, h_n = gh.chunk(3, 1)
    r, u = torch.sigmoid(i_r + h_r), torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    z = (1 - u) * n + u * z
    x = F.linear(z, sd[p + '.output.weight'], sd[p + '.output.bias']) + x
    out.append(x)
  return torch.stack(out, dim=1)


def planner_mlp(sd, fused, target_point, ego_vel, command, cfg, training=False, taps=None):
  """model.py:306-322,359-376 with transformer_decoder_join = False: (B, num_features) globally pooled features ++
  extra-sensor embedding -> MLP join -> GRUCell heads + target-speed MLP.  Returns (checkpoint, speed logits, wp)."""
  vel = F.batch_norm(ego_vel, sd['velocity_normalization.running_mean'].clone(),
                     sd['velocity_normalization.running_var'].clone(), None, None, training=training, momentum=0.1,
                     eps=1e-5)
  es = torch.cat([vel, command], dim=1)
  es = F.relu(F.linear(es, sd['extra_sensor_encoder.0.weight'], sd['extra_sensor_encoder.0.bias']))
  es = F.relu(F.linear(es, sd['extra_sensor_encoder.2.weight'], sd['extra_sensor_encoder.2.bias']))
  x = torch.cat((fused, es), dim=1)
  for i in (0, 2, 4):
    x = F.relu(F.linear(x, sd[f'join.{i}.weight'], sd[f'join.{i}.bias']))
  if taps is not None:
    taps['joined'] = x
  hs = cfg['gru_hidden_size']
  pred_wp = pred_cp = pred_ts = None
  if cfg.get('use_wp_gru', False):
    pred_wp = gru_waypoints_transfuser(sd, 'wp_decoder', x, target_point, cfg.get('pred_len', 8), cfg)
  if cfg.get('use_controller_input_prediction', True):
    pred_cp = gru_waypoints_transfuser(sd, 'checkpoint_decoder', x, target_point, cfg['predict_checkpoint_len'], cfg)
    pred_ts = F.linear(F.relu(F.linear(x[:, :hs], sd['target_speed_network.0.weight'], sd['target_speed_network.0.bias'])),
                       sd['target_speed_network.2.weight'], sd['target_speed_network.2.bias'])
  return pred_cp, pred_ts, pred_wp


def forward(sd, rgb, lidar_bev, target_point, ego_vel, command, cfg=None, training=False, taps=None, dropout=None):
  """LidarCenterNet.forward, model.py:279-392, default GlobalConfig (transFuser backbone, decoder join, all aux
  heads).  Returns the reference's 10-tuple."""
  cfg = cfg or DEFAULT_CFG
  sd = {k: v.float() if torch.is_floating_point(v) else v for k, v in sd.items()}
  bs = rgb.shape[0]
  bev_feature_grid, fused, image_feature_grid = backbone_forward(sd, rgb, lidar_bev, cfg, training, taps=taps,
                                                                 dropout=dropout)
  if taps is not None:
    taps['bev_feature_grid'], taps['fused_features'], taps['image_feature_grid'] = (bev_feature_grid, fused,
                                                                                    image_feature_grid)
  pred_wp = None
  if not cfg.get('transformer_decoder_join', True):
    pred_checkpoint, pred_target_speed, pred_wp = planner_mlp(sd, fused, target_point, ego_vel, command, cfg, training, taps)
  else:
    pl = planner(sd, fused, target_point, ego_vel, command, cfg, training, taps, dropout)
    pred_checkpoint, pred_target_speed = pl[0], pl[1]
    pred_wp = pl[2] if len(pl) > 2 else None
  # model.py:372-389
  pred_semantic = perspective_decoder(sd, 'semantic_decoder', image_feature_grid, cfg)
  pred_depth = torch.sigmoid(perspective_decoder(sd, 'depth_decoder', image_feature_grid, cfg)).squeeze(1)
  b = F.relu(F.conv2d(bev_feature_grid, sd['bev_semantic_decoder.0.weight'], sd['bev_semantic_decoder.0.bias'],
                      padding=1))
  b = F.conv2d(b, sd['bev_semantic_decoder.2.weight'], sd['bev_semantic_decoder.2.bias'])
  b = F.interpolate(b, size=(cfg['lidar_resolution_height'], cfg['lidar_resolution_width']), mode='bilinear',
                    align_corners=False)
  pred_bev_semantic = b * sd['valid_bev_pixels']
  pred_bounding_box = center_net_head(sd, 'head', bev_feature_grid)
  return (pred_wp, pred_target_speed, pred_checkpoint, pred_semantic, pred_bev_semantic, pred_depth,
          pred_bounding_box, None, None, None)


# ---------------------------------------------------------------------------------------------------------------
# a14  CenterNet decode  (center_net.py:172-237, gaussian_target.py:186-264)
# ---------------------------------------------------------------------------------------------------------------
def decode_heatmap(center_heatmap_pred, wh_pred, offset_pred, yaw_class_pred, yaw_res_pred, cfg=None):
  cfg = cfg or DEFAULT_CFG
  k, kernel = cfg['top_k_center_keypoints'], cfg['center_net_max_pooling_kernel']
  b, _, fh, fw = center_heatmap_pred.shape
  hr = float(cfg['lidar_resolution_height'] / fh)
  wr = float(cfg['lidar_resolution_width'] / fw)
  hmax = F.max_pool2d(center_heatmap_pred, kernel, stride=1, padding=(kernel - 1) // 2)
  heat = center_heatmap_pred * (hmax == center_heatmap_pred).float()
  scores, inds = torch.topk(heat.reshape(b, -1), k)
  clses = torch.div(inds, fh * fw, rounding_mode='trunc')
  inds = inds % (fh * fw)
  ys = torch.div(inds, fw, rounding_mode='trunc')
  xs = (inds % fw).int().float()

  def gather(feat):
    feat = feat.permute(0, 2, 3, 1).contiguous().view(b, fh * fw, -1)
    return feat.gather(1, inds.unsqueeze(2).repeat(1, 1, feat.shape[2]))

  wh, off, ycls, yres = gather(wh_pred), gather(offset_pred), gather(yaw_class_pred), gather(yaw_res_pred)
  ycls = torch.argmax(ycls, -1)
  yaw = ycls.float() * (2 * np.pi / float(cfg['num_dir_bins'])) + yres.squeeze(2)
  yaw[yaw > np.pi] -= 2 * np.pi
  xs = xs + off[..., 0]
  ys = ys + off[..., 1]
  zeros = torch.zeros_like(yaw)
  boxes = torch.stack([xs, ys, wh[..., 0], wh[..., 1], yaw, zeros, zeros], dim=2)
  boxes = torch.cat((boxes, clses[..., None], scores[..., None]), dim=-1)
  boxes[:, :, 0] *= wr
  boxes[:, :, 1] *= hr
  boxes[:, :, 2] *= wr
  boxes[:, :, 3] *= hr
  return boxes


# ---------------------------------------------------------------------------------------------------------------
# a15  losses  (model.py:394-445, center_net.py:77-123, transfuser_utils.py:341-364)
# ---------------------------------------------------------------------------------------------------------------
def gaussian_focal_loss_sum(pred, target, alpha=2.0, gamma=4.0):
  eps = 1e-12
  pos = target.eq(1)
  neg_w = (1 - target).pow(gamma)
  pos_loss = -(pred + eps).log() * (1 - pred).pow(alpha) * pos
  neg_loss = -(1 - pred + eps).log() * pred.pow(alpha) * neg_w
  return (pos_loss + neg_loss).sum()


def compute_loss(sd, outputs, labels, cfg=None):
  """model.py:394-445 + center_net.py:77-123.  ``labels`` keys follow train.py:797-820."""
  cfg = cfg or DEFAULT_CFG
  _, pred_ts, pred_cp, pred_sem, pred_bev, pred_depth, bb = outputs[:7]
  loss = {}
  loss['loss_target_speed'] = F.cross_entropy(pred_ts, labels['target_speed'],
                                              weight=torch.tensor(cfg['target_speed_weights']))
  loss['loss_checkpoint'] = torch.mean(torch.abs(pred_cp - labels['checkpoint']))
  loss['loss_semantic'] = F.cross_entropy(pred_sem, labels['semantic'])
  valid = sd['valid_bev_pixels'].squeeze(1).int()
  vis = valid * labels['bev_semantic']
  vis = (valid - 1) + vis
  loss['loss_bev_semantic'] = F.cross_entropy(pred_bev, vis.long(), ignore_index=-1)
  loss['loss_depth'] = F.l1_loss(pred_depth, labels['depth'])
  avg = labels['avg_factor'].sum() + torch.finfo(torch.float32).eps
  pw = labels['pixel_weight']
  loss['loss_center_heatmap'] = gaussian_focal_loss_sum(bb[0], labels['center_heatmap']) / avg
  loss['loss_wh'] = (torch.abs(bb[1] - labels['wh']) * pw).sum() / (avg * 2)
  loss['loss_offset'] = (torch.abs(bb[2] - labels['offset']) * pw).sum() / (avg * 2)
  loss['loss_yaw_class'] = (F.cross_entropy(bb[3], labels['yaw_class'], reduction='none') * pw[:, 0]).sum() / avg
  loss['loss_yaw_res'] = (F.smooth_l1_loss(bb[4], labels['yaw_res'], reduction='none') * pw[:, 0:1]).sum() / avg
  if outputs[0] is not None and 'waypoint' in labels:  # use_wp_gru, model.py:401-402
    loss['loss_wp'] = torch.mean(torch.abs(outputs[0] - labels['waypoint']))
  return loss


LOSS_KEYS = ('loss_target_speed', 'loss_checkpoint', 'loss_semantic', 'loss_bev_semantic', 'loss_depth',
             'loss_center_heatmap', 'loss_wh', 'loss_offset', 'loss_yaw_class', 'loss_yaw_res')


def total_loss(loss):
  """train.py:452-456,889-896: the 10 active weights are 1.0 each, normalised to sum 1."""
  return sum(loss[k] for k in LOSS_KEYS) / len(LOSS_KEYS)


# ---------------------------------------------------------------------------------------------------------------
# a12  camera-frustum mask of the BEV grid  (transfuser_utils.py:596-665, model.py:93-101)
# ---------------------------------------------------------------------------------------------------------------
def valid_bev_pixels(min_x=-32, max_x=32, min_y=-32, max_y=32, pixels_per_meter=4.0, min_z=-10, max_z=14,
                     camera_pos=(-1.5, 0.0, 2.0), fov=110, width=1024, height=256):
  """(1,1,256,256) float mask: 1 where any voxel of the BEV column projects inside the camera image.

  Restates create_projection_grid (identity camera rotation, transfuser_utils.py:620-622) and the max over height +
  transpose at model.py:94-97.  fp32 arithmetic in the same order as the reference so the comparisons agree."""
  mpp = 1.0 / pixels_per_meter
  widths = torch.arange(min_x, max_x, mpp) + (mpp * 0.5)
  depths = torch.arange(min_y, max_y, mpp) + (mpp * 0.5)
  heights = torch.arange(min_z, max_z, mpp) + (mpp * 0.5)
  depths, widths, heights = torch.meshgrid(depths, widths, heights, indexing='ij')
  cloud = torch.stack((depths, widths, heights), dim=0)
  _, d, w, h = cloud.shape
  rot = torch.eye(3)
  t = torch.tensor(list(camera_pos)).unsqueeze(1)
  c2 = (rot.T @ cloud.view(3, -1)) - (rot.T @ t)
  c2 = torch.stack((c2[1], c2[2], c2[0]))
  f = width / (2.0 * np.tan(fov * np.pi / 360.0))
  k = torch.from_numpy(np.array([[f, 0.0, width / 2.0], [0.0, f, height / 2.0], [0.0, 0.0, 1.0]])).float()
  c2 = k @ c2
  z = c2[2:3]
  uv = c2[:2] / z
  uv = uv.view(2, d, w, h)
  z = z.view(1, d, w, h)
  ok = (uv[0:1] >= 0.0) & (uv[0:1] < width) & (uv[1:2] >= 0.0) & (uv[1:2] < height) & (z > 0.0)
  valid = ok.float()  # (1, d, w, h)
  vb = torch.max(valid, dim=3)[0].unsqueeze(1)  # (1,1,d,w)
  return torch.transpose(vb, 2, 3).contiguous()
