"""smoke(): one small invocation of the hot path on cuda:0, checked against the CPU oracle.

1. LiDAR pillar scatter (K1) on one 60k-point cloud: bit-exact vs oracle.lidar_to_histogram_features.
2. TransFuser++ forward, B=1, eval mode, golden state: planner outputs vs oracle.forward within the bf16 floor.
3. One training step (forward + fused losses + backward + AdamW) at B=2: finite losses, parameters move.
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
