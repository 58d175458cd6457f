import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from carla_garage_b200 import ops
which = sys.argv[1] if len(sys.argv) > 1 else 'qkv'
if which == 'qkv':
  x = torch.randn(10240, 1512, device='cuda').to(torch.bfloat16); w = torch.randn(4536, 1512, device='cuda').to(torch.bfloat16)
  f = lambda: ops.linear(x, w)
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
for _ in range(4): f()
torch.cuda.synchronize()
