"""Parameter containers of RegNetY-3.2GF with timm-0.6.7 module names / state_dict keys
(``stem.conv.weight``, ``s1.b1.conv1.bn.running_mean``, ``s1.b1.se.fc1.weight``, ``s1.b1.downsample.conv.weight``),
what ``timm.create_model('regnety_032', features_only=True)`` builds at team_code/transfuser.py:25,52-55.

These modules only HOLD parameters (so checkpoints, ``create_optimizer_groups`` name tests at model.py:586-594 and
``SyncBatchNorm.convert_sync_batchnorm`` keep working); the arithmetic is done by carla_garage_b200.engine on the
sm_100a kernels.  Geometry per timm ``RegNetCfg(w0=80, wa=42.63, wm=2.66, group_size=24, depth=21, se_ratio=0.25)``.
"""
import math

from torch import nn

WIDTHS = (72, 216, 576, 1512)
DEPTHS = (2, 5, 13, 1)
GROUP_WIDTH = 24
STEM_WIDTH = 32


class _NoForward(nn.Module):

  def forward(self, *args, **kwargs):  # pylint: disable=unused-argument
    raise RuntimeError(f'{type(self).__name__} is a parameter container; the forward pass runs in '
                       'carla_garage_b200.engine (CUDA kernels), there is no torch fallback')


class ConvNormAct(_NoForward):

  def __init__(self, cin, cout, k, stride=1, groups=1, act=True):
    super().__init__()
    self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, groups=groups, bias=False)
    self.bn = nn.BatchNorm2d(cout, eps=1e-5, momentum=0.1)
    self.act = act
    self.stride = stride


class SEModule(_NoForward):

  def __init__(self, channels, rd_channels):
    super().__init__()
    self.fc1 = nn.Conv2d(channels, rd_channels, 1, bias=True)
    self.fc2 = nn.Conv2d(rd_channels, channels, 1, bias=True)


class Bottleneck(_NoForward):

  def __init__(self, cin, cout, stride):
    super().__init__()
    self.conv1 = ConvNormAct(cin, cout, 1)
    self.conv2 = ConvNormAct(cout, cout, 3, stride=stride, groups=cout // GROUP_WIDTH)
    self.se = SEModule(cout, int(round(cin * 0.25)))
    self.conv3 = ConvNormAct(cout, cout, 1, act=False)
    self.downsample = ConvNormAct(cin, cout, 1, stride=stride, act=False) if (cin != cout or stride != 1) else None
    self.stride = stride


class _FeatureInfo:

  def __init__(self, info):
    self.info = info


class RegNetY032Features(nn.ModuleDict):
  """FeatureListNet surface used by the reference: .items(), .return_layers, .feature_info.info."""

  def __init__(self, in_chans=3, num_stages=4):
    """num_stages < 4: the later stages are never built (bev_encoder.py:35-37,85-87 delete ``s4`` after timm created
    it; feature_info keeps all five entries, as timm's does)."""
    super().__init__()
    self['stem'] = ConvNormAct(in_chans, STEM_WIDTH, 3, stride=2)
    prev = STEM_WIDTH
    info = [dict(num_chs=STEM_WIDTH, reduction=2, module='stem')]
    red = 2
    for i, (w, d) in enumerate(zip(WIDTHS, DEPTHS)):
      stage = nn.Sequential()
      for j in range(d):
        stage.add_module(f'b{j + 1}', Bottleneck(prev, w, 2 if j == 0 else 1))
        prev = w
      if i < num_stages:
        self[f's{i + 1}'] = stage
      red *= 2
      info.append(dict(num_chs=w, reduction=red, module=f's{i + 1}'))
    self.feature_info = _FeatureInfo(info)
    self.return_layers = {'stem': '0', 's1': '1', 's2': '2', 's3': '3', 's4': '4'}
    self.in_chans = in_chans
    # timm init: conv N(0, sqrt(2 / fan_out)), BN 1/0, zero_init_last (conv3.bn.weight = 0)
    for m in self.modules():
      if isinstance(m, nn.Conv2d):
        fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
        m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
        if m.bias is not None:
          m.bias.data.zero_()
    for m in self.modules():
      if isinstance(m, Bottleneck):
        nn.init.zeros_(m.conv3.bn.weight)
