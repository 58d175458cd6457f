"""Micro-benchmark of the HBM-bound feature-map kernels at the RegNet stage shapes (B=32): achieved GB/s against the
algorithmic bytes each kernel has to move.  Not a bench.py value; used to steer kernel optimisation."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from carla_garage_b200 import ops

FLUSH = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')  # > L2

def timeit(fn, iters=10):
  for _ in range(2): fn()
  torch.cuda.synchronize()
  tot = 0.0
  for _ in range(iters):
    FLUSH.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
  return tot / iters

def bf(*s): return torch.randn(*s, device='cuda').to(torch.bfloat16)
def rec(name, ms, bytes_): print(f'{name:52s} {ms*1e3:9.1f} us  {bytes_/ms/1e6:8.1f} GB/s', flush=True)

B = int(os.environ.get('B', 32))
for (h, w, c, tag) in ((128, 512, 72, 's1.b1.conv1'), (64, 256, 72, 's1'), (32, 128, 216, 's2'), (16, 64, 576, 's3'), (8, 32, 1512, 's4')):
  x, r = bf(B, h, w, c), bf(B, h, w, c)
  sc, sh = torch.rand(c, device='cuda') + 0.5, torch.randn(c, device='cuda')
  n = x.numel() * 2
  y = torch.empty_like(x)
  rec(f'scale_shift_act {tag} relu', timeit(lambda: ops.scale_shift_act(x, sc, sh, ops.ACT_RELU, out=y)), 2 * n)
  pool = torch.zeros(B, c, device='cuda')
  rec(f'scale_shift_act {tag} relu+pool', timeit(lambda: ops.scale_shift_act(x, sc, sh, ops.ACT_RELU, out=y, pool_sum=pool)), 2 * n)
  rec(f'scale_shift_act {tag} +res relu', timeit(lambda: ops.scale_shift_act(x, sc, sh, ops.ACT_RELU, res=r, out=y)), 3 * n)
  gate = torch.rand(B, c, device='cuda')
  rec(f'channel_scale {tag}', timeit(lambda: ops.channel_scale(x, gate, out=y)), 2 * n)
  mean, invstd = torch.randn(c, device='cuda'), torch.rand(c, device='cuda') + 0.5
  dg, db = torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
  rec(f'bn_bwd {tag} relu (reduce+apply)', timeit(lambda: ops.bn_bwd(x, r, y, mean, invstd, sc, ops.ACT_RELU, dg, db)), 7 * n)
  rec(f'bn_bwd {tag} relu +dz', timeit(lambda: ops.bn_bwd(x, r, y, mean, invstd, sc, ops.ACT_RELU, dg, db, want_dz=True)), 8 * n)
  rec(f'bn_bwd {tag} linear', timeit(lambda: ops.bn_bwd(x, None, y, mean, invstd, sc, ops.ACT_NONE, dg, db)), 5 * n)
  rd = max(8, c // 4 // 2 * 2)
  w1, b1 = torch.randn(rd, c, device='cuda'), torch.randn(rd, device='cuda')
  w2, b2 = torch.randn(c, rd, device='cuda'), torch.randn(c, device='cuda')
  rec(f'se_gate {tag} rd={rd}', timeit(lambda: ops.se_gate(pool, h * w, w1, b1, w2, b2)), 4.0 * (2 * rd * c + 2 * B * c))
  g_, hid = ops.se_gate(pool, h * w, w1, b1, w2, b2, want_hidden=True)
  dws = [torch.zeros_like(t) for t in (w1, b1, w2, b2)]
  rec(f'se_bwd {tag} (reduce + mlp adjoint)', timeit(lambda: ops.se_bwd(x, r, g_, hid, pool, h * w, w1, w2, *dws)), 2 * n)
for (rows, c, layout, tag) in ((B * 320, 4536, 0, 'qkv bias C=1512'), (B * 320, 6048, 0, 'mlp relu C=1512'), (B * 320, 1512, 2, 'proj f32'),
                               (B * 65536, 64, 0, 'bev decoder 256x256x64'), (B * 4096, 320, 0, 'head 64x64x320')):
  dy = torch.randn(rows, c, device='cuda') if layout == 2 else bf(rows, c)
  y = bf(rows, c)
  dbias = torch.zeros(c, device='cuda')
  act = ops.ACT_RELU if 'relu' in tag or 'decoder' in tag or 'head' in tag else ops.ACT_NONE
  nb = rows * c * ((4 if layout == 2 else 2) + 2 + (2 if act else 0))
  rec(f'act_bwd {tag}', timeit(lambda: ops.act_bwd(dy, y if act else None, act, 1, rows, c, layout=layout, dbias=dbias)), nb)
# small-channel 3x3 convolutions (perspective decoder tail)
for (h, w, cin, cout, tag) in ((256, 1024, 32, 32, 'dec5'), (256, 1024, 32, 8, 'dec6'), (256, 256, 32, 32, 'bev')):
  x = bf(B, h, w, cin)
  wt = bf(cout, 9, cin)
  ms = timeit(lambda: ops.smallc_conv3x3(x, wt), iters=5)
  m = B * h * w
  rec(f'smallc_conv3x3 {tag} {cin}->{cout}  ({2.0*m*cin*cout*9/ms/1e9:.0f} TFLOP/s)', ms, 2.0 * m * (cin + cout))
# RegNet group convs on the haloed-tile kernel (image branch shapes; stride 2 = first block of the stage)
for (h, w, c, stride, tag) in ((128, 512, 72, 2, 's1.b1'), (64, 256, 72, 1, 's1'), (64, 256, 216, 2, 's2.b1'), (32, 128, 216, 1, 's2'),
                               (16, 64, 576, 1, 's3'), (8, 32, 1512, 1, 's4'), (16, 16, 576, 1, 's3 lidar')):
  x = bf(B, h, w, c)
  wp = ops.pack_gconv_halo(torch.randn(c, 24, 3, 3, device='cuda'))
  st = (torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda'))
  ms = timeit(lambda: ops.gconv3x3(x, wp, stride, stats=st), iters=5)
  m = B * (h // stride) * (w // stride)
  rec(f'gconv3x3+stats {tag} {h}x{w} C={c} s{stride}  ({2.0*m*c*24*9/ms/1e9:.0f} TFLOP/s)', ms, 2.0 * (B * h * w * c + m * c))
