"""GPU parity of the whole TransFuser++ forward (C-ABI kernels under the reference's module surface) against
(a) golden vectors made from the unmodified reference (tests/golden/make_golden.py) and (b) the CPU oracle run live.
Tolerance: north_star's bf16 bound is 1e-2 relative; that bound is enforced per component on exact inputs
(test_blocks_in_isolation: every RegNet block / fusion block / head given the oracle's fp32 input).  End to end, a
randomly initialised TransFuser++ amplifies bf16 STORAGE rounding by ~2.5x per stage (measured with
oracle/bf16_emulation.py: an ideal implementation that only rounds conv/linear operands and results to bf16 is already
2-4e-2 away from fp32 at the last stage), so the end-to-end assertion is "no worse than 2.5x the recorded bf16 floor
of that tensor, and never worse than 1e-2 where the floor allows it" — see DESIGN.md "Numerics"."""
import os

import numpy as np
import pytest
import torch

from tests.golden.sampling import sample

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
TOL = 1e-2


def rel(a, b):
  a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
  return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope='module')
def net(oracle_state):
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  from carla_garage_b200.config import GlobalConfig
  from carla_garage_b200.nn import LidarCenterNet
  m = LidarCenterNet(GlobalConfig())
  m.load_state_dict(oracle_state, strict=True)
  return m.cuda()


def _inputs(b, seed):
  from carla_garage_b200 import synth
  return {k: v.cuda() for k, v in synth.make_inputs(b, seed=seed).items()}


@pytest.mark.noisy
def test_forward_eval_vs_golden(net):
  from carla_garage_b200 import ops
  g = np.load(os.path.join(GOLDEN, 'forward_eval_b2.npz'))
  net.eval()
  net.engine.debug_taps = {}
  with torch.no_grad():
    out = net(**_inputs(2, 11))
  torch.cuda.synchronize()
  taps = net.engine.debug_taps
  net.engine.debug_taps = None
  report = {}
  for k, t in taps.items():
    if t is None or ('tap_' + k) not in g.files:
      continue
    t = ops.nhwc_to_nchw(t) if t.dim() == 4 else t
    report[k] = rel(sample(t), g['tap_' + k])
  print('\n'.join(f'  tap {k}: {v:.2e}' for k, v in report.items()))
  errs = {
      'pred_target_speed': rel(out[1], g['pred_target_speed']),
      'pred_checkpoint': rel(out[2], g['pred_checkpoint']),
      'pred_semantic': rel(sample(out[3]), g['pred_semantic']),
      'pred_bev_semantic': rel(sample(out[4]), g['pred_bev_semantic']),
      'pred_depth': rel(sample(out[5]), g['pred_depth']),
  }
  for n, t in zip(('heatmap', 'wh', 'offset', 'yaw_class', 'yaw_res'), out[6][:5]):
    errs['bb_' + n] = rel(sample(t), g['bb_' + n])
  print('\n'.join(f'  out {k}: {v:.2e}' for k, v in errs.items()))
  assert out[3].shape == (2, 7, 256, 1024) and out[4].shape == (2, 11, 256, 256) and out[5].shape == (2, 256, 1024)
  assert out[0] is None and out[7] is None and out[6][5] is None
  for k, v in {**report, **errs}.items():
    floor = float(g['bf16floor_' + k]) if ('bf16floor_' + k) in g.files else 0.0
    assert v < max(TOL, 2.5 * floor), (k, v, floor)
  # full-size norms (not just the sub-sampled fixtures)
  for n, t in (('pred_semantic', out[3]), ('pred_bev_semantic', out[4]), ('pred_depth', out[5])):
    assert abs(float(t.norm()) / float(g['norm_' + n]) - 1) < 5 * TOL
  # decode: boxes with a clear score margin must agree with the reference's decode of ITS heat maps
  boxes = net.head.get_bboxes(*out[6])
  assert boxes.shape == (2, 100, 9)
  assert rel(boxes[..., 8], g['boxes'][..., 8]) < max(TOL, 2.5 * float(g['bf16floor_bb_heatmap']))


@pytest.mark.noisy
def test_forward_eval_vs_live_oracle(net, oracle_state):
  from carla_garage_b200 import synth
  from oracle import tfpp_oracle as orc
  inp = synth.make_inputs(1, seed=21)
  net.eval()
  with torch.no_grad():
    got = net(**{k: v.cuda() for k, v in inp.items()})
    torch.set_num_threads(min(os.cpu_count() or 1, 16))
    want = orc.forward(oracle_state, **inp)
  g = np.load(os.path.join(GOLDEN, 'forward_eval_b2.npz'))
  names = {1: 'pred_target_speed', 2: 'pred_checkpoint', 3: 'pred_semantic', 4: 'pred_bev_semantic', 5: 'pred_depth'}
  for i in (1, 2, 3, 4, 5):
    assert rel(got[i], want[i]) < max(TOL, 3 * float(g['bf16floor_' + names[i]])), i
  for n, a, b in zip(('heatmap', 'wh', 'offset', 'yaw_class', 'yaw_res'), got[6][:5], want[6][:5]):
    assert rel(a, b) < max(TOL, 3 * float(g['bf16floor_bb_' + n]))


@pytest.mark.noisy
def test_forward_train_mode_vs_golden(net, oracle_state):
  """Training-mode forward: BatchNorm batch statistics (+ running-stat update), dropout off."""
  g = np.load(os.path.join(GOLDEN, 'train_b2.npz'))
  net.load_state_dict(oracle_state, strict=True)
  net.train()
  with torch.no_grad():
    out = net(**_inputs(2, 11))
  ge = np.load(os.path.join(GOLDEN, 'forward_eval_b2.npz'))
  assert rel(out[1], g['pred_target_speed']) < max(TOL, 3 * float(ge['bf16floor_pred_target_speed']))
  assert rel(out[2], g['pred_checkpoint']) < max(TOL, 3 * float(ge['bf16floor_pred_checkpoint']))
  assert rel(sample(out[3]), g['pred_semantic']) < max(TOL, 3 * float(ge['bf16floor_pred_semantic']))
  bn = net.backbone.image_encoder['stem'].bn
  assert int(bn.num_batches_tracked) == 1
  assert not torch.equal(bn.running_mean.cpu(), oracle_state['backbone.image_encoder.stem.bn.running_mean'])
  net.load_state_dict(oracle_state, strict=True)
  net.eval()


def test_backbone_module_api(net):
  """TransfuserBackbone.forward keeps the reference contract (NCHW f32 features)."""
  net.eval()
  inp = _inputs(1, 5)
  with torch.no_grad():
    feats, fused, grid = net.backbone(inp['rgb'], inp['lidar_bev'])
  assert feats.shape == (1, 64, 64, 64) and fused.shape == (1, 1512, 8, 8) and grid.shape == (1, 1512, 8, 32)
  assert feats.dtype == torch.float32


def test_blocks_in_isolation(net, oracle_state):
  """north_star's 1e-2 bf16 bound, enforced per component: every RegNet block and every fusion block gets the
  ORACLE's fp32 input and must reproduce the oracle's output within 1e-2 relative L2."""
  from carla_garage_b200 import ops, synth
  from oracle import tfpp_oracle as orc
  net.eval()
  eng, sd = net.engine, oracle_state
  inp = synth.make_inputs(2, seed=11)
  to_dev = lambda t: ops.nchw_to_nhwc(t.cuda().contiguous())
  torch.set_num_threads(min(os.cpu_count() or 1, 16))
  worst = 0.0
  with torch.no_grad():
    x_img = orc._conv_bn(sd, 'backbone.image_encoder.stem', orc.normalize_imagenet(inp['rgb']), False, stride=2)
    x_lid = orc._conv_bn(sd, 'backbone.lidar_encoder.stem', inp['lidar_bev'], False, stride=2)
    assert rel(ops.nhwc_to_nchw(eng.stem(inp['rgb'].cuda(), net.backbone.image_encoder['stem'], False, True)), x_img) < TOL
    assert rel(ops.nhwc_to_nchw(eng.stem(inp['lidar_bev'].cuda(), net.backbone.lidar_encoder['stem'], False, False)),
               x_lid) < TOL
    for i in range(4):
      for name, enc in (('image', net.backbone.image_encoder), ('lidar', net.backbone.lidar_encoder)):
        x = x_img if name == 'image' else x_lid
        for j, blk in enumerate(enc[f's{i + 1}']):
          want = orc.regnet_block(sd, f'backbone.{name}_encoder.s{i + 1}.b{j + 1}', x, False, stride=2 if j == 0 else 1)
          e = rel(ops.nhwc_to_nchw(eng.regnet_block(to_dev(x), blk, False)), want)
          worst = max(worst, e)
          assert e < TOL, (name, i, j, e)
          x = want
        if name == 'image':
          x_img = x
        else:
          x_lid = x
      w_img, w_lid = orc.fuse_features(sd, 'backbone', x_img, x_lid, i, orc.DEFAULT_CFG)
      g_img, g_lid = eng.fuse(to_dev(x_img), to_dev(x_lid), i, False)
      assert rel(ops.nhwc_to_nchw(g_img), w_img) < TOL and rel(ops.nhwc_to_nchw(g_lid), w_lid) < TOL
      x_img, x_lid = w_img, w_lid
  print('worst isolated block error', worst)
