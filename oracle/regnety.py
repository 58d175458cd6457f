"""ORACLE (test infrastructure, never imported by the product path).

Plain-torch CPU restatement of the one piece of arithmetic on the TransFuser++ path that is NOT in the
reference tree: ``timm==0.6.7`` (``team_code/requirements.txt:163``) ``regnety_032`` built with
``features_only=True`` as used at ``team_code/transfuser.py:25,52-55``.  timm is absent from /root/reference and
from this image, so this follows timm 0.6.7's published definition:

* ``timm/models/regnet.py``: ``model_cfgs['regnety_032'] = RegNetCfg(w0=80, wa=42.63, wm=2.66, group_size=24,
  depth=21, se_ratio=0.25)``; ``generate_regnet`` + ``adjust_widths_groups_comp`` give widths [72,216,576,1512],
  depths [2,5,13,1], groups [3,9,24,63]; stem = ConvNormAct(in,32,k3,s2); ``Bottleneck`` (bottle_ratio 1):
  conv1 1x1+BN+ReLU -> conv2 3x3 grouped (stride in first block of a stage)+BN+ReLU -> SE(rd=round(0.25*in_chs))
  -> conv3 1x1+BN -> +shortcut (downsample = 1x1 stride-s conv+BN when shape changes) -> ReLU.
* ``timm/models/features.py``: ``FeatureListNet`` is an ``nn.ModuleDict`` of exactly the modules up to the last
  feature (stem,s1..s4), with ``return_layers`` and ``feature_info.info``.
* ``timm/models/layers/squeeze_excite.py``: ``SEModule``: mean over (H,W) -> fc1 (1x1 conv, bias) -> ReLU -> fc2 ->
  sigmoid gate.

Parity status: UNPINNED by the reference (it ships no tests/golden vectors for team_code); the
geometry is cross-checked by parameter count (17.92 M for in_chans=3) and MACs (3.2 GMAC @224²) in
tests/test_oracle.py.  State-dict keys follow timm (``stem.conv.weight``, ``s1.b1.conv1.bn.running_mean``,
``s1.b1.se.fc1.weight``, ``s1.b1.downsample.conv.weight`` ...), consistent with the name tests at
``team_code/model.py:586-594``.
"""
import numpy as np
import torch
from torch import nn


def regnet_widths(w0=80, wa=42.63, wm=2.66, depth=21, group_size=24, q=8):
  """timm 0.6.7 regnet.py generate_regnet + adjust_widths_groups_comp (bottle_ratio = 1)."""
  widths_cont = np.arange(depth) * wa + w0
  width_exps = np.round(np.log(widths_cont / w0) / np.log(wm))
  widths = w0 * np.power(wm, width_exps)
  widths = (np.round(np.divide(widths, q)) * q).astype(int)
  stage_widths, stage_depths = np.unique(widths, return_counts=True)
  stage_widths, stage_depths = stage_widths.tolist(), stage_depths.tolist()
  groups = [min(group_size, w) for w in stage_widths]
  # quantise widths to a multiple of the group width
  stage_widths = [int(round(w / g) * g) for w, g in zip(stage_widths, groups)]
  return stage_widths, stage_depths, groups


class ConvNormAct(nn.Module):
  """timm ConvBnAct: .conv (no bias) + .bn (BatchNormAct2d: BatchNorm2d with optional fused ReLU)."""

  def __init__(self, cin, cout, k, stride=1, groups=1, act=True):
    super().__init__()
    self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, groups=groups, bias=False)
    self.bn = nn.BatchNorm2d(cout, eps=1e-5, momentum=0.1)
    self.act = act

  def forward(self, x):
    x = self.bn(self.conv(x))
    return torch.relu(x) if self.act else x


class SEModule(nn.Module):

  def __init__(self, channels, rd_channels):
    super().__init__()
    self.fc1 = nn.Conv2d(channels, rd_channels, 1, bias=True)
    self.fc2 = nn.Conv2d(rd_channels, channels, 1, bias=True)

  def forward(self, x):
    s = x.mean((2, 3), keepdim=True)
    s = self.fc2(torch.relu(self.fc1(s)))
    return x * torch.sigmoid(s)


class Bottleneck(nn.Module):
  """RegNet Y block, bottle_ratio = 1."""

  def __init__(self, cin, cout, stride, group_size, se_ratio=0.25):
    super().__init__()
    groups = cout // group_size
    self.conv1 = ConvNormAct(cin, cout, 1)
    self.conv2 = ConvNormAct(cout, cout, 3, stride=stride, groups=groups)
    self.se = SEModule(cout, int(round(cin * se_ratio)))
    self.conv3 = ConvNormAct(cout, cout, 1, act=False)
    if cin != cout or stride != 1:
      self.downsample = ConvNormAct(cin, cout, 1, stride=stride, act=False)
    else:
      self.downsample = None

  def forward(self, x):
    shortcut = x if self.downsample is None else self.downsample(x)
    x = self.conv3(self.se(self.conv2(self.conv1(x))))
    return torch.relu(x + shortcut)


class _FeatureInfo:

  def __init__(self, info):
    self.info = info


class RegNetYFeatures(nn.ModuleDict):
  """What ``timm.create_model('regnety_032', features_only=True)`` returns (a FeatureListNet)."""

  def __init__(self, in_chans=3, zero_init_last=True):
    super().__init__()
    widths, depths, groups = regnet_widths()
    assert widths == [72, 216, 576, 1512] and depths == [2, 5, 13, 1], (widths, depths)
    self['stem'] = ConvNormAct(in_chans, 32, 3, stride=2)
    prev = 32
    info = [dict(num_chs=32, reduction=2, module='stem')]
    red = 2
    for i, (w, d, g) in enumerate(zip(widths, depths, groups)):
      blocks = nn.Sequential()
      for j in range(d):
        blocks.add_module(f'b{j + 1}', Bottleneck(prev, w, 2 if j == 0 else 1, g))
        prev = w
      self[f's{i + 1}'] = blocks
      red *= 2
      info.append(dict(num_chs=w, reduction=red, module=f's{i + 1}'))
    self.feature_info = _FeatureInfo(info)
    self.return_layers = {'stem': '0', 's1': '1', 's2': '2', 's3': '3', 's4': '4'}
    # timm init: conv N(0, sqrt(2/fan_out)), BN 1/0, zero_init_last -> conv3.bn.weight = 0
    for m in self.modules():
      if isinstance(m, nn.Conv2d):
        fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
        m.weight.data.normal_(0, (2.0 / fan_out)**0.5)
        if m.bias is not None:
          m.bias.data.zero_()
    if zero_init_last:
      for m in self.modules():
        if isinstance(m, Bottleneck):
          nn.init.zeros_(m.conv3.bn.weight)

  def forward(self, x):
    out = []
    for _, m in self.items():
      x = m(x)
      out.append(x)
    return out


def timm_factory(name, pretrained=False, features_only=True, in_chans=3):
  """Drop-in for timm.create_model as called at team_code/transfuser.py:25,52-55."""
  del pretrained  # ImageNet weights are an external download; tests share an explicit state_dict instead
  if name != 'regnety_032' or not features_only:
    raise NotImplementedError(f'oracle timm shim only restates regnety_032 features_only (asked: {name})')
  return RegNetYFeatures(in_chans=in_chans)
