"""CPU tests: the oracle restatement against the golden vectors made from the unmodified reference
(tests/golden/make_golden.py), and — in the build container only — against the live reference modules."""
import os

import numpy as np
import pytest
import torch

from carla_garage_b200 import compat, synth
from oracle import tfpp_oracle as orc
from oracle import regnety
from tests.golden.sampling import sample

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def rel(a, b):
  a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
  return float((a - b).norm() / (b.norm() + 1e-30))


def test_regnety_geometry():
  widths, depths, groups = regnety.regnet_widths()
  assert widths == [72, 216, 576, 1512] and depths == [2, 5, 13, 1] and groups == [24] * 4
  m = regnety.RegNetYFeatures(3)
  n = sum(p.numel() for p in m.parameters())
  assert n == 17923338  # 17.92 M
  assert [i['num_chs'] for i in m.feature_info.info] == [32, 72, 216, 576, 1512]
  y = m(torch.zeros(1, 3, 64, 64))
  assert [t.shape[1:] for t in y] == [(32, 32, 32), (72, 16, 16), (216, 8, 8), (576, 4, 4), (1512, 2, 2)]


def test_pillar_scatter_oracle_vs_golden():
  g = np.load(os.path.join(GOLDEN, 'pillar_scatter.npz'))
  pts = synth.make_point_clouds(2, seed=7).numpy()
  clouds = dict(cloud0=pts[0], cloud1=pts[1], edge=g['edge_points'], empty=np.zeros((0, 3), np.float32))
  for name, cloud in clouds.items():
    for gp in (0, 1):
      out = orc.lidar_to_histogram_features(cloud, bool(gp))
      want = g[f'{name}_gp{gp}'].astype(np.float32) / 5.0
      assert out.shape == want.shape and out.dtype == np.float32
      assert np.array_equal(out, want), (name, gp)
  # survey-recorded known answers of the reference function
  rng = np.random.default_rng(0)
  cloud = np.stack([rng.uniform(-40, 40, 60000), rng.uniform(-40, 40, 60000), rng.uniform(-1, 4, 60000)], 1)
  out = orc.lidar_to_histogram_features(cloud, False)
  assert abs(float(out.sum()) - 5831.2002) < 1e-2 and int((out > 0).sum()) == 23606 and out.max() == 1.0


def test_valid_bev_pixels_vs_golden():
  want = np.load(os.path.join(GOLDEN, 'valid_bev_pixels.npz'))['valid']
  got = orc.valid_bev_pixels().numpy()
  assert got.shape == (1, 1, 256, 256)
  assert np.array_equal(got.astype(np.uint8), want)


def test_forward_eval_vs_golden(oracle_state):
  g = np.load(os.path.join(GOLDEN, 'forward_eval_b2.npz'))
  inp = synth.make_inputs(2, seed=11)
  taps = {}
  torch.set_num_threads(min(os.cpu_count() or 1, 16))
  with torch.no_grad():
    out = orc.forward(oracle_state, **inp, taps=taps)
  assert rel(out[1], g['pred_target_speed']) < 1e-4
  assert rel(out[2], g['pred_checkpoint']) < 1e-4
  assert rel(sample(out[3]), g['pred_semantic']) < 1e-4
  assert rel(sample(out[4]), g['pred_bev_semantic']) < 1e-4
  assert rel(sample(out[5]), g['pred_depth']) < 1e-4
  for n, t in zip(('heatmap', 'wh', 'offset', 'yaw_class', 'yaw_res'), out[6][:5]):
    assert rel(sample(t), g['bb_' + n]) < 1e-4, n
    assert abs(float(t.norm()) / float(g['norm_bb_' + n]) - 1) < 1e-4
  for k in ('img_stem', 'lid_stem', 'img_s1_pre', 'lid_s4_pre', 'bev_feature_grid', 'fused_features',
            'image_feature_grid', 'joined'):
    assert rel(sample(taps[k]), g['tap_' + k]) < 1e-4, k
  boxes = orc.decode_heatmap(*out[6][:5])
  assert boxes.shape == (2, 100, 9)
  assert rel(boxes, g['boxes']) < 1e-4


def test_train_losses_vs_golden(oracle_state):
  g = np.load(os.path.join(GOLDEN, 'train_b2.npz'))
  inp = synth.make_inputs(2, seed=11)
  lab = synth.make_labels(2, seed=13)
  sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v)
        for k, v in oracle_state.items()}
  torch.set_num_threads(min(os.cpu_count() or 1, 16))
  out = orc.forward(sd, **inp, training=True)
  # forward() detaches; run the differentiable path explicitly
  loss = orc.compute_loss(oracle_state, out, lab)
  for k in orc.LOSS_KEYS:
    assert abs(float(loss[k]) - float(g[k])) <= 2e-4 * max(1.0, abs(float(g[k]))), k
  assert abs(float(orc.total_loss(loss)) - float(g['total'])) < 2e-4


@pytest.mark.reference
@pytest.mark.skipif(not compat.reference_available(), reason='needs /root/reference (build container)')
def test_oracle_vs_live_reference(oracle_state):
  compat.install(regnety.timm_factory)
  from config import GlobalConfig  # pylint: disable=import-outside-toplevel
  from model import LidarCenterNet  # pylint: disable=import-outside-toplevel
  net = LidarCenterNet(GlobalConfig()).eval()
  net.load_state_dict(oracle_state, strict=True)
  inp = synth.make_inputs(1, seed=5)
  with torch.no_grad():
    want = net(**inp)
    got = orc.forward(oracle_state, **inp)
  for i in (1, 2, 3, 4, 5):
    assert rel(got[i], want[i]) < 1e-5, i
  for a, b in zip(got[6][:5], want[6][:5]):
    assert rel(a, b) < 1e-5
  pts = synth.make_point_clouds(1, seed=3).numpy()[0]
  for gp in (False, True):
    assert np.array_equal(orc.lidar_to_histogram_features(pts, gp), net.data.lidar_to_histogram_features(pts, gp))


def test_philox_known_answers():
  """oracle/philox.py against the Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors)."""
  from oracle import philox as ph
  kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
         ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
         ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
          (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
  for ctr, key, want in kat:
    got = ph.philox4x32_10([ctr[0]], [ctr[1]], [ctr[2]], [ctr[3]], key[0], key[1])
    assert tuple(int(w[0]) for w in got) == want
  m = ph.multiplier((3, 1000), 0.1, 1234, 1, 3)
  assert abs(float((m == 0).float().mean()) - 0.1) < 0.02 and abs(float(m.max()) - 1 / 0.9) < 1e-6
  import torch
  t = torch.ones(2, 8)
  s1, s2 = ph.DropoutStream(5, 1), ph.DropoutStream(5, 1)
  assert torch.equal(s1(t, 0.5), s2(t, 0.5)) and s1.site == 1 and torch.equal(s1(t, 0.0), t) and s1.site == 1


def test_nms_oracle_closed_form_iou():
  """oracle/nms.py (transfuser_utils.py:409-452 with shapely's polygon IoU restated as convex clipping) against
  closed-form intersections."""
  import math
  from oracle import nms
  a = (0.0, 0.0, 2.0, 1.0, 0.0)
  assert abs(nms.iou_bbs(a, a) - 1.0) < 1e-12
  assert abs(nms.iou_bbs(a, (1.0, 0.0, 2.0, 1.0, 0.0)) - 0.6) < 1e-12          # 3x2 overlap of two 4x2 boxes
  assert nms.iou_bbs(a, (10.0, 0.0, 2.0, 1.0, 0.3)) == 0.0
  sq, sq45 = (3.0, -2.0, 1.0, 1.0, 0.0), (3.0, -2.0, 1.0, 1.0, math.pi / 4)
  octagon = 8.0 * (math.sqrt(2.0) - 1.0)                                          # square ∩ its 45 degree copy
  assert abs(nms.iou_bbs(sq, sq45) - octagon / (8.0 - octagon)) < 1e-12
  # a quarter turn swaps the extents: 2x1 vs 1x2 half extents overlap in the central 2x2 square
  assert abs(nms.iou_bbs(a, (0.0, 0.0, 2.0, 1.0, math.pi / 2)) - 4.0 / 12.0) < 1e-12
  # IoU is symmetric and invariant under a common rigid motion
  b1, b2 = (1.0, 2.0, 2.5, 1.2, 0.4), (2.0, 2.5, 1.5, 2.2, -0.9)
  i12 = nms.iou_bbs(b1, b2)
  assert abs(i12 - nms.iou_bbs(b2, b1)) < 1e-12 and 0.0 < i12 < 1.0
  th, c, s = 0.7, math.cos(0.7), math.sin(0.7)
  mv = lambda b: (c * b[0] - s * b[1] + 5.0, s * b[0] + c * b[1] - 3.0, b[2], b[3], b[4] + th)
  assert abs(nms.iou_bbs(mv(b1), mv(b2)) - i12) < 1e-12


def test_nms_oracle_greedy_order_and_vehicle_frame():
  from oracle import nms
  boxes = [[(0.0, 0.0, 2.0, 1.0, 0.0, 0.9), (0.5, 0.0, 2.0, 1.0, 0.0, 0.8)],      # member 0: second overlaps the first
           [(10.0, 0.0, 2.0, 1.0, 0.0, 0.7), (0.1, 0.1, 2.0, 1.0, 0.1, 0.95)],    # member 1: best box of all + a far one
           None]                                                                  # member without detections
  kept = nms.non_maximum_suppression(boxes, 0.2)
  assert [float(k[-1]) for k in kept] == [0.95, 0.7]
  assert nms.non_maximum_suppression([[], None], 0.2) == []
  v = nms.bb_image_to_vehicle_system([140.0, 100.0, 8.0, 4.0, 0.3, 1.0, 0.0, 0.0, 0.9], 4.0, -32.0, -32.0)
  assert np.allclose(v[:5], [(100.0 - 128.0) / 4, (140.0 - 128.0) / 4, 1.0, 2.0, -0.3])


def test_targets_oracle_vs_reference_golden():
  """oracle/targets.py == the unmodified reference rasteriser (data.py:698-791, gaussian_target.py) on 8 cases incl.
  borders, an empty sample and two boxes sharing a centre pixel (tests/golden/make_targets_golden.py)."""
  from oracle import targets
  g = np.load(os.path.join(GOLDEN, 'targets.npz'))
  n = sum(1 for k in g.files if k.startswith('boxes'))
  assert n == 8
  for i in range(n):
    t, avg = targets.get_targets(g[f'boxes{i}'])
    assert avg == int(g[f'avg{i}']), i
    for k, v in t.items():
      want = g[f'{k}{i}']
      if v.dtype.kind == 'i':
        assert np.array_equal(v, want), (i, k)
      else:
        assert np.allclose(v, want, rtol=0, atol=1e-6), (i, k, float(np.abs(v - want).max()))
    assert np.array_equal(t['center_heatmap_target'] == 1, g[f'center_heatmap_target{i}'] == 1)


def test_oracle_mlp_join_vs_reference_golden():
  """The oracle's restatement of the original TransFuser planner (transformer_decoder_join = False, use_wp_gru:
  model.py:184-209,359-376,870-913; transfuser.py:188-197) against the unmodified reference: training-mode forward,
  the eleven losses and the gradients of every planner parameter (tests/golden/make_golden_mlp_join.py)."""
  import json
  from carla_garage_b200 import synth
  from oracle import tfpp_oracle as orc
  g = np.load(os.path.join(GOLDEN, 'mlp_join_b2.npz'))
  keys = json.load(open(os.path.join(GOLDEN, 'mlp_join_keys.json')))
  sd = synth.mlp_join_state(GOLDEN)
  sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k and not k.startswith(('valid_bev', 'loss_'))
            else v.clone()) for k, v in sd.items()}
  assert list(sd.keys()) == sorted(keys['order']) or set(sd) == set(keys['order'])
  cfg = dict(orc.DEFAULT_CFG, transformer_decoder_join=False, use_wp_gru=True, pred_len=8)
  inp, lab = synth.make_inputs(2, seed=11), synth.make_labels(2, seed=13)
  lab['waypoint'] = synth.make_waypoint_labels(2, 8, seed=13)
  out = orc.forward(sd, **inp, cfg=cfg, training=True)
  for k, i in (('pred_wp', 0), ('pred_target_speed', 1), ('pred_checkpoint', 2)):
    assert float((out[i] - torch.from_numpy(g[k])).norm() / torch.from_numpy(g[k]).norm()) < 2e-4, k
  losses = orc.compute_loss(sd, out, lab, cfg)
  assert len(losses) == 11
  for k, v in losses.items():
    assert abs(float(v) - float(g[k])) <= 2e-4 * max(1.0, abs(float(g[k]))), k
  (sum(losses.values()) / len(losses)).backward()
  n = 0
  for key in g.files:
    if key.startswith('grad_'):
      name = key[5:]
      got = sd[name].grad.flatten()[:512]
      want = torch.from_numpy(g[key])
      assert float((got - want).norm() / (want.norm() + 1e-30)) < 2e-3, name
      n += 1
  assert n >= 20


def test_align_transforms_reproduce_reference_align():
  """carla_garage_b200.dataio.align_transforms (host side of tfpp_pillar_scatter_aligned) applied in float64 + the
  oracle histogram == CARLA_Data.align + lidar_to_histogram_features of the unmodified reference
  (tests/golden/make_align_golden.py)."""
  from carla_garage_b200 import dataio, synth
  from oracle import tfpp_oracle as orc
  g = np.load(os.path.join(GOLDEN, 'align.npz'))
  pts = synth.make_point_clouds(len(g['cases']), seed=21, n_points=20000).numpy()
  for i, c in enumerate(g['cases']):
    xf = dataio.align_transforms({'pos_global': (c[0], c[1]), 'theta': c[2]}, {'pos_global': (c[3], c[4]), 'theta': c[5]},
                                 y_augmentation=c[6], yaw_augmentation=c[7])
    p = pts[i].astype(np.float64)
    for tx, ty, tz, yaw in xf:
      rot = np.array([[np.cos(yaw), -np.sin(yaw), 0.0], [np.sin(yaw), np.cos(yaw), 0.0], [0.0, 0.0, 1.0]])
      p = (rot.T @ (p - np.array([tx, ty, tz])).T).T
    for gp in (0, 1):
      h = orc.lidar_to_histogram_features(p, bool(gp))
      assert np.array_equal(np.round(h * 5).astype(np.uint8), g[f'hist{i}_gp{gp}']), (i, gp)
