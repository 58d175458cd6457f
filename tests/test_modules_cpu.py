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


def test_pack_plan_index_maps_reproduce_every_gather_pack():
  """engine.PackPlan learns 'packed element -> flat parameter index' by running the pack code on index tensors; a
  plain gather through that map must reproduce the torch-built pack bit for bit (host logic of tfpp_gather_pack)."""
  import torch
  from carla_garage_b200 import engine as E
  g = torch.Generator().manual_seed(0)
  shapes = {'conv': (48, 16, 3, 3), 'gconv': (96, 24, 3, 3), 'lin_a': (40, 24), 'lin_b': (16, 24), 'c1': (24, 8, 1, 1),
            'c2': (12, 16, 1, 1), 'pos': (1, 5, 24), 'vec': (40,)}
  total = sum(int(torch.tensor(s).prod()) for s in shapes.values())
  flat = torch.randn(total, generator=g)
  params, off = {}, 0
  for k, s in shapes.items():
    n = int(torch.tensor(s).prod())
    params[k] = flat[off:off + n].view(s)
    off += n
  plan = E.PackPlan(flat)
  cases = [('conv', (params['conv'],), ()), ('conv_t', (params['conv'],), ()), ('conv_rows_pad', (params['conv'],), (64,)),
           ('conv_dgrad_smallc', (params['conv'],), (64,)), ('gconv', (params['gconv'],), ()),
           ('gconv_t', (params['gconv'],), ()), ('gconv_halo', (params['gconv'],), ()),
           ('gconv_halo_t', (params['gconv'],), ()), ('gconv_halo_umma', (params['gconv'],), ()),
           ('gconv_halo_umma_t', (params['gconv'],), ()), ('conv_halo_umma', (params['conv'],), (64,)),
           ('conv_halo_umma_t', (params['conv'],), (48,)), ('linear', (params['lin_a'],), ()), ('linear_t', (params['lin_a'],), ()),
           ('rows', (params['lin_a'],), (8, 24)), ('rows_t', (params['lin_a'],), (8, 24)),
           ('cols', (params['lin_a'],), (8, 24)), ('cols_t', (params['lin_a'],), (3, 24)),
           ('rows_f32', (params['lin_a'],), (8, 24)), ('cat_linear', (params['lin_a'], params['lin_b']), ()),
           ('cat_linear_t', (params['lin_a'], params['lin_b']), ()), ('cat_rows', (params['lin_a'], params['lin_b']), (0, 8)),
           ('cat_rows_f32', (params['lin_a'], params['lin_b']), (0, 8)), ('cat_f32', (params['vec'], params['lin_b']), ()),
           ('cat_conv', (params['c1'], params['c1']), ()), ('cat_conv_t', (params['c1'], params['c1']), ()),
           ('blockdiag_1x1', (params['c1'], params['c2']), ()), ('blockdiag_1x1_t', (params['c1'], params['c2']), ()),
           ('conv_cin', (params['conv'],), (8, 16)), ('conv_cin_t', (params['conv'],), (0, 8)),
           ('conv_cin_pad', (params['conv'],), (24,)), ('conv_cin_pad_t', (params['conv'],), (24,)),
           ('repeat_rows', (params['pos'],), (3,))]
  assert {c[0] for c in cases} == set(E._GATHER_KINDS)  # pylint: disable=protected-access
  for kind, ps, extra in cases:
    out = E._build_pack(kind, ps, extra)  # pylint: disable=protected-access
    key = (kind,) + tuple(id(p) for p in ps) + extra
    plan.register(key, kind, ps, extra, out)
    assert key in plan.pending, kind
    _, items, is_tuple, _, _ = plan.pending[key]
    assert is_tuple == isinstance(out, tuple)
    for idx, like in items:
      idx = idx.long()
      got = torch.where(idx >= 0, flat[idx.clamp(min=0)], torch.zeros(())).to(like.dtype).view(like.shape)
      assert torch.equal(got, like), kind
  # a tensor outside the flat buffer is left to the ordinary cache
  other = torch.randn(8, 8)
  plan.register(('linear', id(other)), 'linear', (other,), (), E._build_pack('linear', (other,), ()))  # pylint: disable=protected-access
  assert ('linear', id(other)) not in plan.pending


def test_halo_umma_layout_math_reproduces_conv3x3():
  """Host-side model of the experimental halo-UMMA kernel (csrc/halo_umma.cu): channel-chunk-major halo planes with a
  64-pixel pitch, tap = start-address shift of (ky*64 + kx) pixels, weights packed [tap][k chunk][n][8].  Emulating
  exactly that addressing in torch must reproduce F.conv2d; junk columns (x = 62, 63) are dropped."""
  import torch
  import torch.nn.functional as F
  from carla_garage_b200 import ops
  g = torch.Generator().manual_seed(0)
  cin, cout, n_pad, h, w = 16, 7, 16, 11, 70
  x = torch.randn(1, h, w, cin, generator=g)
  wt = torch.randn(cout, cin, 3, 3, generator=g) * 0.1
  want = F.conv2d(x.permute(0, 3, 1, 2), wt, padding=1)[0]  # (cout, h, w)
  wp = ops.pack_halo_umma_weight(wt, n_pad, dt=torch.float32)  # (9, K/8, N, 8)
  PW, THO = 64, 8
  PH = THO + 2
  got = torch.zeros(cout, h, w)
  for y0 in range(0, h, THO):
    for x0 in range(0, w, PW - 2):
      # what the TMA box {8, 64, 10, 1} at (c*8, x0-1, y0-1) writes, with zero fill outside the image
      planes = torch.zeros(cin // 8, PH * PW + 8, 8)
      for py in range(PH):
        for px in range(PW):
          yy, xx = y0 - 1 + py, x0 - 1 + px
          if 0 <= yy < h and 0 <= xx < w:
            planes[:, py * PW + px, :] = x[0, yy, xx].view(cin // 8, 8)
      acc = torch.zeros(THO * PW, n_pad)
      m = torch.arange(THO * PW)
      for tap in range(9):
        ky, kx = divmod(tap, 3)
        a = planes[:, m + ky * PW + kx, :]                       # (K/8, M, 8): rows = linear pixels shifted by the tap
        acc += torch.einsum('cmj,cnj->mn', a, wp[tap])           # contraction over (chunk, 8)
      for mm in range(THO * PW):
        ty, tx = divmod(mm, PW)
        if tx < PW - 2 and y0 + ty < h and x0 + tx < w:
          got[:, y0 + ty, x0 + tx] = acc[mm, :cout]
  assert torch.allclose(got, want, atol=1e-4), float((got - want).abs().max())
  # the input-gradient pack: conv of dY with it equals autograd's dX
  xg = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
  wt2 = torch.randn(16, cin, 3, 3, generator=g) * 0.1
  dy = torch.randn(1, 16, h, w, generator=g)
  F.conv2d(xg, wt2, padding=1).backward(dy)
  wpt = ops.pack_halo_umma_weight(wt2, cin, transpose=True, dt=torch.float32)  # (9, 16/8, cin, 8)
  w_eq = wpt.permute(2, 1, 3, 0).reshape(cin, 16, 3, 3)  # back to (out = cin, in = 16, ky, kx) of the equivalent conv
  assert torch.allclose(F.conv2d(dy, w_eq, padding=1), xg.grad, atol=1e-4)


def test_halo_umma_gconv_layout_math_reproduces_group_conv():
  """Host-side model of the experimental grouped halo-UMMA kernel: a 72-channel slab in ten 8-channel planes (pitch 64),
  group g contracts planes 3g..3g+3 (K = 32, the last 8 channels meet zero weights) against its [tap][4][32][8] weight
  block and keeps 24 of 32 output columns."""
  import torch
  import torch.nn.functional as F
  from carla_garage_b200 import ops
  g = torch.Generator().manual_seed(1)
  c, h, w = 144, 6, 70
  x = torch.randn(1, h, w, c, generator=g)
  wt = torch.randn(c, 24, 3, 3, generator=g) * 0.1
  want = F.conv2d(x.permute(0, 3, 1, 2), wt, padding=1, groups=c // 24)[0]
  wp = ops.pack_halo_gconv_weight(wt, dt=torch.float32)  # (C/24, 9, 4, 32, 8)
  assert wp.shape == (c // 24, 9, 4, 32, 8)
  PW, THO, NPL = 64, 4, 10
  PH = THO + 2
  got = torch.zeros(c, h, w)
  for slab in range(c // 72):
    for y0 in range(0, h, THO):
      for x0 in range(0, w, PW - 2):
        planes = torch.zeros(NPL, PH * PW + 8, 8)
        for pl in range(NPL):
          ch0 = slab * 72 + pl * 8
          for py in range(PH):
            for px in range(PW):
              yy, xx = y0 - 1 + py, x0 - 1 + px
              if 0 <= yy < h and 0 <= xx < w and ch0 < c:   # TMA zero fill outside the tensor (also beyond C)
                planes[pl, py * PW + px, :] = x[0, yy, xx, ch0:ch0 + 8]
        m = torch.arange(THO * PW)
        for grp in range(3):
          acc = torch.zeros(THO * PW, 32)
          for tap in range(9):
            ky, kx = divmod(tap, 3)
            a = planes[3 * grp:3 * grp + 4][:, m + ky * PW + kx, :]            # (4 chunks, M, 8)
            acc += torch.einsum('cmj,cnj->mn', a, wp[slab * 3 + grp, tap])
          for mm in range(THO * PW):
            ty, tx = divmod(mm, PW)
            if tx < PW - 2 and y0 + ty < h and x0 + tx < w:
              c0 = slab * 72 + grp * 24
              got[c0:c0 + 24, y0 + ty, x0 + tx] = acc[mm, :24]
  assert torch.allclose(got, want, atol=1e-4), float((got - want).abs().max())
  # the input-gradient pack is the pack of the equivalent forward conv (in/out swapped inside each group, taps flipped)
  xg = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
  dy = torch.randn(1, c, h, w, generator=g)
  F.conv2d(xg, wt, padding=1, groups=c // 24).backward(dy)
  wpt = ops.pack_halo_gconv_weight(wt, transpose=True, dt=torch.float32)
  w_eq = wpt[:, :, :3, :24, :].permute(0, 3, 2, 4, 1).reshape(c // 24, 24, 24, 3, 3).reshape(c, 24, 3, 3)
  assert torch.allclose(F.conv2d(dy, w_eq, padding=1, groups=c // 24), xg.grad, atol=1e-4)


def test_ensemble_reduction_matches_sensor_agent_semantics():
  """inference.ensemble_outputs = sensor_agent.py:481-483,527-531: mean over the members of the softmaxed target-speed
  logits and of the predicted checkpoints."""
  import torch
  from carla_garage_b200.inference import ensemble_outputs
  g = torch.Generator().manual_seed(3)
  outs = []
  for _ in range(3):
    o = [None] * 10
    o[1] = torch.randn(4, 4, generator=g)
    o[2] = torch.randn(4, 10, 2, generator=g)
    outs.append(tuple(o))
  probs, cps = ensemble_outputs(outs)
  want_p = sum(torch.softmax(o[1], dim=1) for o in outs) / 3
  want_c = sum(o[2] for o in outs) / 3
  assert torch.allclose(probs, want_p, atol=1e-6) and torch.allclose(cps, want_c, atol=1e-6)
  assert torch.allclose(probs.sum(1), torch.ones(4), atol=1e-6)


def test_mlp_join_config_matches_reference_keys():
  """transformer_decoder_join = False + use_wp_gru (the original TransFuser planner, model.py:184-209): same state_dict
  keys, shapes and registration ORDER as the unmodified reference (tests/golden/make_golden_mlp_join.py)."""
  from carla_garage_b200.nn import LidarCenterNet
  ref = json.load(open(os.path.join(GOLDEN, 'mlp_join_keys.json')))
  cfg = GlobalConfig()
  cfg.transformer_decoder_join = False
  cfg.use_wp_gru = True
  sd = LidarCenterNet(cfg).state_dict()
  assert list(sd.keys()) == ref['order']
  for k, v in sd.items():
    assert list(v.shape) == ref['shapes'][k], k
