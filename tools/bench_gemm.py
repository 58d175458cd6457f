"""Micro-benchmark of the tcgen05 kernels on the layer shapes of TransFuser++ (B=32).  Prints achieved TFLOP/s and the
activation GB/s per shape.  Not a bench.py value; used to steer kernel optimisation."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from carla_garage_b200 import ops

def timeit(fn, iters=20):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters

def bf(*s): return torch.randn(*s, device='cuda').to(torch.bfloat16)

B = int(os.environ.get('B', 32))
rows = []
def rec(name, ms, flop, bytes_):
  rows.append((name, ms, flop / ms / 1e9, bytes_ / ms / 1e6))
  print(f'{name:46s} {ms*1e3:9.1f} us  {flop/ms/1e9:8.1f} TFLOP/s  {bytes_/ms/1e6:8.1f} GB/s', flush=True)

only = sys.argv[1] if len(sys.argv) > 1 else ''
# linears of the C=1512 fusion block
for (m, k, n, tag) in ((B*320, 1512, 4536, 'qkv C=1512'), (B*320, 1512, 6048, 'mlp1 C=1512'), (B*320, 6048, 1512, 'mlp2 C=1512'),
                       (B*320, 576, 1728, 'qkv C=576'), (B*320, 576, 2304, 'mlp1 C=576'), (B*320, 216, 648, 'qkv C=216'),
                       (B*11, 256, 2048, 'dec ff1'), (B*64, 1512, 256, 'change_channel')):
  if only and only not in tag: continue
  x, w = bf(m, k), bf(n, k)
  ms = timeit(lambda: ops.linear(x, w))
  rec(f'linear {tag} M={m} K={k} N={n}', ms, 2.0*m*n*k, 2.0*(m*k + n*k + m*n))
# 1x1 convs of the RegNet stages (image branch) with BN statistics epilogue
for (h, w_, cin, cout, tag) in ((64, 256, 72, 72, 's1'), (32, 128, 216, 216, 's2'), (16, 64, 576, 576, 's3'), (8, 32, 1512, 1512, 's4'),
                                (128, 512, 32, 72, 's1.b1.conv1')):
  if only and only not in tag and only != 'conv1x1': continue
  x, wt = bf(B, h, w_, cin), bf(cout, 1, cin)
  st = (torch.zeros(cout, device='cuda'), torch.zeros(cout, device='cuda'))
  ms = timeit(lambda: ops.conv_gemm(x, wt, stats=st))
  m = B*h*w_
  rec(f'conv1x1+stats {tag} {h}x{w_} {cin}->{cout}', ms, 2.0*m*cin*cout, 2.0*(m*cin + m*cout))
# grouped 3x3
for (h, w_, c, tag) in ((64, 256, 72, 's1'), (32, 128, 216, 's2'), (16, 64, 576, 's3')):
  if only and only != 'grouped': continue
  x = bf(B, h, w_, c)
  wp = ops.pack_grouped_conv_weight(torch.randn(c, 24, 3, 3, device='cuda'))
  ms = timeit(lambda: ops.conv_gemm(x, wp, taps=ops.TAPS_3X3, k_per_tile=48, a_c_per_ntile=48, bn=48))
  m = B*h*w_
  rec(f'grouped3x3 {tag} {h}x{w_} C={c}', ms, 2.0*m*c*24*9, 2.0*(2*m*c))
# dense 3x3 of the perspective decoders
for (h, w_, cin, cout, tag) in ((8, 32, 1512, 128, 'dec1'), (64, 256, 64, 32, 'dec3'), (64, 256, 32, 32, 'dec4'), (256, 1024, 32, 32, 'dec5'),
                                (256, 1024, 32, 7, 'dec6'), (64, 64, 64, 320, 'head')):
  if only and only != 'conv3x3' and only not in tag: continue
  x, wt = bf(B, h, w_, cin), bf(cout, 9, cin)
  ms = timeit(lambda: ops.conv_gemm(x, wt, taps=ops.TAPS_3X3), iters=5)
  m = B*h*w_
  rec(f'conv3x3 {tag} {h}x{w_} {cin}->{cout}', ms, 2.0*m*cin*cout*9, 2.0*(m*cin + m*cout))
# wgrad
for (h, w_, cin, cout, k, tag) in ((256, 1024, 32, 32, 3, 'dec5'), (64, 256, 72, 72, 1, 's1'), (16, 64, 576, 576, 1, 's3'), (1, B*320, 1512, 4536, 1, 'qkv'),
                                   (1, B*320, 6048, 1512, 1, 'mlp2')):
  if only and only != 'wgrad': continue
  bb = B if h > 1 else 1
  x, dy = bf(bb, h, w_, cin), bf(bb, h, w_, cout)
  taps = ops.TAPS_3X3 if k == 3 else ops.TAPS_1X1
  out = torch.zeros(cout, k*k, cin, device='cuda')
  ms = timeit(lambda: ops.conv_wgrad(dy, x, taps=taps, out=out, out_strides=(k*k*cin, cin, 1)), iters=5)
  m = bb*h*w_
  rec(f'wgrad {tag} {h}x{w_} {cin}->{cout} k{k}', ms, 2.0*m*cin*cout*k*k, 2.0*(m*cin + m*cout))
json.dump(rows, open(os.path.join(ROOT, 'gpurun_out', 'bench_gemm.json'), 'w'))
