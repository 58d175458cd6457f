"""CPU tests of the host side: module tree / state_dict compatibility with the reference, C-ABI symbol export,
config defaults, tile heuristics.  No kernel is launched here."""
import ctypes
import json
import os
import re

import pytest
import torch

from carla_garage_b200 import _lib, compat, ops
from carla_garage_b200.config import GlobalConfig

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
ROOT = os.path.dirname(os.path.dirname(__file__))


def test_state_dict_matches_reference_keys():
  from carla_garage_b200.nn import LidarCenterNet
  ref = json.load(open(os.path.join(GOLDEN, 'state_dict_keys.json')))
  net = LidarCenterNet(GlobalConfig())
  sd = net.state_dict()
  assert set(sd.keys()) == set(ref['shapes'].keys())
  for k, v in sd.items():
    assert list(v.shape) == ref['shapes'][k], k
    assert str(v.dtype) == ref['dtypes'][k], k
  assert sum(p.numel() for p in net.parameters()) == 120351026


def test_valid_bev_pixels_matches_reference():
  import numpy as np
  from carla_garage_b200.nn.model import valid_bev_pixels
  want = np.load(os.path.join(GOLDEN, 'valid_bev_pixels.npz'))['valid']
  assert np.array_equal(valid_bev_pixels(GlobalConfig()).numpy().astype(np.uint8), want)


def test_optimizer_groups_cover_all_parameters():
  from carla_garage_b200.nn import LidarCenterNet
  net = LidarCenterNet(GlobalConfig())
  groups = net.create_optimizer_groups(0.01)
  assert sum(len(g['params']) for g in groups) == len(list(net.parameters()))


def test_unknown_backbone_raises_like_reference():
  from carla_garage_b200.nn import LidarCenterNet
  cfg = GlobalConfig()
  cfg.backbone = 'nope'
  with pytest.raises(ValueError):
    LidarCenterNet(cfg)


def test_library_exports_every_declared_symbol():
  """libtfpp.so loads without a GPU and exports everything include/tfpp.h declares."""
  header = open(os.path.join(ROOT, 'include', 'tfpp.h')).read()
  declared = set(re.findall(r'\b(tfpp_[a-z0-9_]+)\s*\(', header))
  declared.discard('tfpp_conv_gemm_args')
  lib = ctypes.CDLL(_lib.LIB_PATH)
  for name in declared:
    assert hasattr(lib, name), name
  assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
  assert _lib.load().tfpp_abi_version() >= 1


def test_cpu_tensor_is_rejected_not_silently_computed():
  with pytest.raises(RuntimeError):
    ops.pillar_scatter(torch.zeros(1, 4, 3))


def test_tile_and_bn_heuristics():
  for h, w in ((1, 352), (8, 8), (8, 32), (16, 16), (16, 64), (64, 64), (64, 256), (256, 1024), (1, 64), (32, 128)):
    tw, th, nb = ops.pick_tile(h, w)
    assert tw * th * nb == 128 and tw >= min(w, 128)
  assert ops.pick_bn(72) == 80 and ops.pick_bn(1512) == 256 and ops.pick_bn(6048) == 256 and ops.pick_bn(7) == 16
  taps = ops.taps_3x3_stride2(3)
  assert len(taps) == 9 and taps[4] == (0, 0, 0, 4) and taps[0] == (-1, -1, 9, 0)


@pytest.mark.reference
@pytest.mark.skipif(not compat.reference_available(), reason='needs /root/reference')
def test_config_defaults_match_reference():
  from oracle.regnety import timm_factory
  compat.install(timm_factory)
  from config import GlobalConfig as RefConfig  # pylint: disable=import-outside-toplevel
  ref, mine = RefConfig(), GlobalConfig()
  for k, v in vars(mine).items():
    assert hasattr(ref, k), k
    assert getattr(ref, k) == v, (k, getattr(ref, k), v)
