"""smoke(): one small invocation of the hot path on cuda:0, checked against the CPU oracle.

1. LiDAR pillar scatter (K1) on one 60k-point cloud: bit-exact vs oracle.lidar_to_histogram_features.
2. TransFuser++ forward, B=1, eval mode, golden state: planner outputs vs oracle.forward within the bf16 floor, and the
   same forward in the fp32 parity mode within north_star's 1e-3.
3. Agent-side kernels: CenterNet decode + rotated-IoU NMS vs oracle/nms.py; CenterNet target rasteriser vs oracle/targets.py.
4. One training step (forward + fused losses + backward + AdamW) at B=2: finite losses, parameters move.
"""
import os

import numpy as np
import torch


def run():
  if not torch.cuda.is_available():
    raise RuntimeError('smoke() needs a CUDA device')
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  golden = os.path.join(root, 'tests', 'golden')
  from . import ops, synth
  from .config import GlobalConfig
  from .nn import LidarCenterNet
  from .training import Trainer
  from oracle import tfpp_oracle as orc  # checker only
  torch.cuda.set_device(0)
  pts = synth.make_point_clouds(1, seed=3)
  got = ops.pillar_scatter(pts.cuda(), use_ground_plane=True).cpu().numpy()[0]
  want = orc.lidar_to_histogram_features(pts[0].numpy(), True)
  assert np.array_equal(got, want), 'pillar scatter differs from the oracle'
  sd = synth.golden_state(golden)
  net = LidarCenterNet(GlobalConfig())
  net.load_state_dict(sd, strict=True)
  net = net.cuda().eval()
  inp = synth.make_inputs(1, seed=21)
  with torch.no_grad():
    out = net(**{k: v.cuda() for k, v in inp.items()})
    torch.set_num_threads(min(os.cpu_count() or 1, 16))  # many-core hosts: torch CPU kernels collapse when oversubscribed
    ref = orc.forward(sd, **inp)
  g = np.load(os.path.join(golden, 'forward_eval_b2.npz'))

  def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))

  for i, name in ((1, 'pred_target_speed'), (2, 'pred_checkpoint'), (3, 'pred_semantic')):
    e, floor = rel(out[i], ref[i]), float(g['bf16floor_' + name])
    assert e < max(1e-2, 3 * floor), (name, e, floor)
  with ops.precision('fp32'), torch.no_grad():   # same schedule, fp32 storage + contractions: north_star's 1e-3
    out32 = net(**{k: v.cuda() for k, v in inp.items()})
  for i in (1, 2, 3, 4, 5):
    assert rel(out32[i], ref[i]) < 1e-3, ('fp32 mode', i, rel(out32[i], ref[i]))
  for a, b in zip(out32[6][:5], ref[6][:5]):
    assert rel(a, b) < 1e-3
  # agent side: decode + NMS of the detections, label rasteriser
  from oracle import nms as onms, targets as otargets  # checkers only
  dec = net.head.get_bboxes(*out32[6])
  kept, count = ops.nms_rotated(dec, 0.05, 0.2, to_vehicle=True)
  want_boxes = onms.ensemble_boxes([dec[0].cpu().numpy()], 0.05, 0.2)
  assert int(count[0]) == len(want_boxes), (int(count[0]), len(want_boxes))
  if want_boxes:
    assert np.allclose(kept[0, :len(want_boxes)].cpu().numpy(), np.stack(want_boxes), rtol=1e-4, atol=1e-4)
  boxes = synth.make_gt_boxes(1, seed=9)[0]
  lab_dev = ops.centernet_targets(torch.from_numpy(boxes)[None].cuda(), torch.tensor([len(boxes)], dtype=torch.int32).cuda())
  want_t, want_avg = otargets.get_targets(boxes)
  assert np.allclose(lab_dev['center_heatmap'][0].cpu().numpy(), want_t['center_heatmap_target'], atol=2e-6)
  assert float(lab_dev['avg_factor'][0]) == float(want_avg)
  net.train()
  tr = Trainer(net)
  before = tr.st.flat[:1000].clone()
  inp2 = {k: v.cuda() for k, v in synth.make_inputs(2, seed=5).items()}
  lab2 = {k: v.cuda().contiguous() for k, v in synth.make_labels(2, seed=6).items()}
  _, losses = tr.step(inp2, lab2)
  torch.cuda.synchronize()
  vals = {k: float(v) for k, v in losses.items()}
  assert all(np.isfinite(v) for v in vals.values()), vals
  assert not torch.equal(before, tr.st.flat[:1000]), 'optimizer step did not change the parameters'
  print('smoke ok', {k: round(v, 4) for k, v in vals.items()})
