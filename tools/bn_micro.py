"""Per-shape timing of the BatchNorm passes of the train step (scale_shift_act forward, bn_bwd = reduce + apply) at the
B=32 shapes: 'cold' = L2 flushed before every call, 'warm' = 20 back-to-back calls.  TFPP_BN_STREAM=0 selects the
previous kernels."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from carla_garage_b200 import ops

SHAPES = {'i0': (32, 128, 512, 72), 'i1': (32, 64, 256, 72), 'i2': (32, 32, 128, 216), 'i3': (32, 16, 64, 576),
          'i4': (32, 8, 32, 1512), 'l2': (32, 32, 32, 216), 'l3': (32, 16, 16, 576)}
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')


def timeit(f):
  for _ in range(3): f()
  torch.cuda.synchronize()
  ts = []
  for _ in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
  ts.sort()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(20): f()
  e1.record(); torch.cuda.synchronize()
  return ts[len(ts) // 2], e0.elapsed_time(e1) * 1e3 / 20


tot = {'ssa': [0, 0], 'bnbwd': [0, 0]}
for name, shp in SHAPES.items():
  for gated in (False, True):
    c = shp[3]
    raw = torch.randn(shp, device='cuda').to(torch.bfloat16); dy = torch.randn(shp, device='cuda').to(torch.bfloat16)
    mean, invstd = torch.randn(c, device='cuda') * 0.1, torch.rand(c, device='cuda') + 0.5
    gamma = torch.rand(c, device='cuda') + 0.5
    fs, fh = gamma * invstd, -mean * gamma * invstd
    gate = torch.rand(shp[0], c, device='cuda') if gated else None
    pg = torch.randn(shp[0], c, device='cuda') * 1e-3 if gated else None
    dg, db = torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
    out = torch.empty_like(raw)
    pool = torch.zeros(shp[0], c, device='cuda') if gated else None
    mb = raw.numel() * 2 / 1e6
    for op, f, passes in (('ssa', lambda: ops.scale_shift_act(raw, fs, fh, ops.ACT_RELU, out=out, pool_sum=pool), 2),
                          ('bnbwd', lambda: ops.bn_bwd(dy, None, raw, mean, invstd, gamma, ops.ACT_RELU, dg, db, gate=gate,
                                                       pool_grad=pg, fwd_affine=(fs, fh)), 5)):
      cold, warm = timeit(f)
      tot[op][0] += cold; tot[op][1] += warm
      print(f'{op:6s} {name}{"g" if gated else " "} {mb:7.1f} MB/tensor  cold {cold:7.1f} us ({passes * mb / cold:5.2f} TB/s)'
            f'  warm {warm:7.1f} us ({passes * mb / warm:5.2f} TB/s)')
print('sum', {k: [round(x, 1) for x in v] for k, v in tot.items()}, 'TFPP_BN_STREAM =', os.environ.get('TFPP_BN_STREAM', '1'))
