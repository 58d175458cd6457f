"""One small eager training step (B=2, dropout on, three streams) + the agent-side kernels, for compute-sanitizer:
  compute-sanitizer --tool memcheck python tools/one_step.py
  PYTORCH_NO_CUDA_MEMORY_CACHING=1 python tools/one_step.py        (allocator stress: every free is a real cudaFree)
Prints the ten losses of the step; exits non-zero if any is not finite."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from carla_garage_b200 import ops, synth  # noqa: E402
from carla_garage_b200.config import GlobalConfig  # noqa: E402
from carla_garage_b200.nn import LidarCenterNet  # noqa: E402
from carla_garage_b200.training import Trainer  # noqa: E402

torch.manual_seed(0)
net = LidarCenterNet(GlobalConfig())
net.load_state_dict(synth.golden_state(os.path.join(ROOT, 'tests', 'golden')), strict=True)
tr = Trainer(net.cuda().train(), use_optim_groups=True)
b = 2
inp = {k: v.cuda() for k, v in synth.make_inputs(b, seed=3).items()}
inp['lidar_bev'] = ops.pillar_scatter(synth.make_point_clouds(b, seed=3).cuda())
lab = {k: v.cuda().contiguous() for k, v in synth.make_labels(b, seed=4).items()}
boxes = synth.make_gt_boxes(b, seed=5)
bx = torch.zeros(b, 30, 8)
cnt = torch.zeros(b, dtype=torch.int32)
for i, c in enumerate(boxes):
  bx[i, :len(c)] = torch.from_numpy(c)
  cnt[i] = len(c)
lab.update({k: v for k, v in ops.centernet_targets(bx.cuda(), cnt.cuda()).items() if k in lab})   # labels rasterised on the device
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
first = None
for _ in range(steps):
  out, losses = tr.step(inp, lab)
  if first is None:
    first = {k: float(v) for k, v in losses.items()}
torch.cuda.synchronize()
vals = {k: float(v) for k, v in losses.items()}
print('first step', first)
print(vals)
net.eval()
with torch.no_grad():
  o = net(**inp)
  dec = net.head.get_bboxes(*o[6])
  kept, count = ops.nms_rotated(torch.cat([dec, dec], dim=1).contiguous(), 0.05, 0.2, to_vehicle=True)
torch.cuda.synchronize()
print('boxes kept', count.tolist())
sys.exit(0 if all(v == v and abs(v) < 1e6 for v in vals.values()) else 1)
