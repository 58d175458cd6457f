import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from carla_garage_b200 import ops
which = sys.argv[1] if len(sys.argv) > 1 else 'qkv'
if which == 'qkv':
  x = torch.randn(10240, 1512, device='cuda').to(torch.bfloat16); w = torch.randn(4536, 1512, device='cuda').to(torch.bfloat16)
  f = lambda: ops.linear(x, w)
elif which == 'c576':  # RegNet stage-3 1x1 conv with BatchNorm statistics (the most frequent GEMM of the step)
  x = torch.randn(32, 16, 64, 576, device='cuda').to(torch.bfloat16); w = torch.randn(576, 1, 576, device='cuda').to(torch.bfloat16)
  st = (torch.zeros(576, device='cuda'), torch.zeros(576, device='cuda'))
  f = lambda: ops.conv_gemm(x, w, stats=st)
elif which == 'mlp':  # fusion MLP up-projection at scale 4: M=10240, K=1512, N=6048
  x = torch.randn(10240, 1512, device='cuda').to(torch.bfloat16); w = torch.randn(6048, 1512, device='cuda').to(torch.bfloat16)
  f = lambda: ops.linear(x, w)
elif which == 'bevlift':
  from carla_garage_b200.config import GlobalConfig
  from carla_garage_b200.nn.bev_encoder import lift_tables, projection_grid
  grid, ok = projection_grid(GlobalConfig())
  norm = torch.finfo(torch.float32).eps + ok.sum(3).unsqueeze(1)
  vbp = torch.transpose(ok.max(3)[0].unsqueeze(1), 2, 3).contiguous()
  tables = tuple(t.cuda() for t in lift_tables(grid, norm, vbp, 32, 128))
  img = torch.randn(32, 32, 128, 32, device='cuda').to(torch.bfloat16)
  f = lambda: ops.bev_lift(img, tables, 256, 256)
elif which == 'dec5':
  x = torch.randn(32, 256, 1024, 32, device='cuda').to(torch.bfloat16); w = torch.randn(32, 9, 32, device='cuda').to(torch.bfloat16)
  f = lambda: ops.conv_gemm(x, w, taps=ops.TAPS_3X3)
elif which == 's1stats':
  x = torch.randn(32, 64, 256, 72, device='cuda').to(torch.bfloat16); w = torch.randn(72, 1, 72, device='cuda').to(torch.bfloat16)
  st = (torch.zeros(72, device='cuda'), torch.zeros(72, device='cuda'))
  f = lambda: ops.conv_gemm(x, w, stats=st)
elif which == 'smallc':
  x = torch.randn(32, 256, 1024, 32, device='cuda').to(torch.bfloat16); w = torch.randn(32, 9, 32, device='cuda').to(torch.bfloat16)
  f = lambda: ops.smallc_conv3x3(x, w)
elif which == 'grouped':
  x = torch.randn(32, 64, 256, 72, device='cuda').to(torch.bfloat16)
  wp = ops.pack_grouped_conv_weight(torch.randn(72, 24, 3, 3, device='cuda'))
  f = lambda: ops.conv_gemm(x, wp, taps=ops.TAPS_3X3, k_per_tile=48, a_c_per_ntile=48, bn=48)
elif which == 'gconv':
  x = torch.randn(32, 64, 256, 72, device='cuda').to(torch.bfloat16)
  wp = ops.pack_gconv_halo(torch.randn(72, 24, 3, 3, device='cuda'))
  st = (torch.zeros(72, device='cuda'), torch.zeros(72, device='cuda'))
  f = lambda: ops.gconv3x3(x, wp, 1, stats=st)
elif which == 'ssa':
  x = torch.randn(32, 64, 256, 72, device='cuda').to(torch.bfloat16); y = torch.empty_like(x)
  sc, sh = torch.rand(72, device='cuda'), torch.rand(72, device='cuda')
  f = lambda: ops.scale_shift_act(x, sc, sh, ops.ACT_RELU, out=y)
elif which == 'halo':
  x = torch.randn(32, 256, 1024, 32, device='cuda').to(torch.bfloat16)
  wp = ops.pack_halo_umma_weight(torch.randn(32, 32, 3, 3, device='cuda'), 32)
  f = lambda: ops.halo_conv3x3(x, wp)
elif which.startswith('attn'):
  c = int(which[4:] or 576)
  qkv = torch.randn(32 * 320, 3 * c, device='cuda').to(torch.bfloat16)
  f = lambda: ops.fusion_attn(qkv, 32, 320, c, 4)
elif which.startswith('battn'):
  c = int(which[5:] or 576)
  qkv = torch.randn(32 * 320, 3 * c, device='cuda').to(torch.bfloat16); do = torch.randn(32 * 320, c, device='cuda').to(torch.bfloat16)
  f = lambda: ops.fusion_attn_bwd(qkv, do, 32, 320, c, 4)
elif which == 'pillar':
  from carla_garage_b200 import synth
  pts = synth.make_point_clouds(32, seed=1).cuda()
  f = lambda: ops.pillar_scatter(pts)
elif which == 'nms':
  bx = torch.rand(64, 300, 9, device='cuda'); bx[..., :2] = bx[..., :2] * 200 + 28; bx[..., 2:4] = bx[..., 2:4] * 8 + 3
  f = lambda: ops.nms_rotated(bx, 0.3, 0.2, to_vehicle=True)
elif which == 'targets':
  from carla_garage_b200 import synth
  import numpy as np
  cs = synth.make_gt_boxes(32, seed=2)
  bx = torch.zeros(32, 30, 8); cnt = torch.zeros(32, dtype=torch.int32)
  for i, c_ in enumerate(cs): bx[i, :len(c_)] = torch.from_numpy(c_); cnt[i] = len(c_)
  bx, cnt = bx.cuda(), cnt.cuda()
  f = lambda: ops.centernet_targets(bx, cnt)
elif which.startswith('bnbwd_') or which.startswith('ssa_'):
  # BatchNorm backward / forward apply at the step's shapes: i1/i2/i3 = image stages, l3 = LiDAR stage 3, g = SE-gated
  shp = {'i1': (32, 64, 256, 72), 'i2': (32, 32, 128, 216), 'i3': (32, 16, 64, 576), 'l3': (32, 16, 16, 576),
         'i0': (32, 128, 512, 72)}[which.split('_')[1][:2]]
  gated = which.endswith('g')
  c = shp[3]
  raw = torch.randn(shp, device='cuda').to(torch.bfloat16); dy = torch.randn(shp, device='cuda').to(torch.bfloat16)
  mean, invstd, gamma = torch.randn(c, device='cuda') * 0.1, torch.rand(c, device='cuda') + 0.5, torch.rand(c, device='cuda') + 0.5
  fs, fh = gamma * invstd, -mean * gamma * invstd
  gate = torch.rand(shp[0], c, device='cuda') if gated else None
  pg = torch.randn(shp[0], c, device='cuda') * 1e-3 if gated else None
  dg, db = torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
  out = torch.empty_like(raw)
  if which.startswith('bnbwd_'):
    f = lambda: ops.bn_bwd(dy, None, raw, mean, invstd, gamma, ops.ACT_RELU, dg, db, gate=gate, pool_grad=pg, fwd_affine=(fs, fh))
  else:
    pool = torch.zeros(shp[0], c, device='cuda') if gated else None
    f = lambda: ops.scale_shift_act(raw, fs, fh, ops.ACT_RELU, out=out, pool_sum=pool)
for _ in range(4): f()
torch.cuda.synchronize()
if len(sys.argv) > 2 and sys.argv[2] == 'warm':  # back-to-back launches, operands as warm in L2 as they fit
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(20): f()
  e1.record(); torch.cuda.synchronize()
  print(f'{which}: warm {e0.elapsed_time(e1) * 1e3 / 20:.1f} us per call')
elif len(sys.argv) > 2 and sys.argv[2] == 'time':  # CUDA-event timing, L2 flushed between launches
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
  ts = []
  for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
  ts.sort()
  print(f'{which}: median {ts[len(ts) // 2]:.1f} us  min {ts[0]:.1f} us')
else:
  f()
  torch.cuda.synchronize()
