"""GPU parity tests of the individual sm_100a kernels, called through the C ABI (carla_garage_b200.ops), against
fp32 torch restatements of the same op evaluated on the bf16-rounded operands the kernel sees."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(scope='module')
def ops():
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  from carla_garage_b200 import ops as o
  return o


def rel(a, b):
  a, b = a.double().cpu(), b.double().cpu()
  return float((a - b).norm() / (b.norm() + 1e-30))


def rnd(*shape, seed=0, scale=1.0):
  g = torch.Generator().manual_seed(seed)
  return (torch.randn(*shape, generator=g) * scale).cuda()


def bf(x):
  return x.to(torch.bfloat16)


# ------------------------------------------------------------------------------------------------ K1
def test_pillar_scatter_bit_exact(ops):
  from carla_garage_b200 import synth
  from oracle import tfpp_oracle as orc
  g = np.load(os.path.join(GOLDEN, 'pillar_scatter.npz'))
  pts = synth.make_point_clouds(2, seed=7)
  for gp in (0, 1):
    out = ops.pillar_scatter(pts.cuda(), use_ground_plane=bool(gp)).cpu().numpy()
    for b in range(2):
      want = g[f'cloud{b}_gp{gp}'].astype(np.float32) / 5.0
      assert np.array_equal(out[b], want)
      assert np.array_equal(out[b], orc.lidar_to_histogram_features(pts[b].numpy(), bool(gp)))
    edge = torch.from_numpy(g['edge_points']).cuda()[None]
    out = ops.pillar_scatter(edge, use_ground_plane=bool(gp)).cpu().numpy()[0]
    assert np.array_equal(out, g[f'edge_gp{gp}'].astype(np.float32) / 5.0)
    empty = torch.zeros((1, 0, 3), device='cuda')
    out = ops.pillar_scatter(empty, use_ground_plane=bool(gp)).cpu().numpy()
    assert out.shape == (1, 1 + gp, 256, 256) and not out.any()
  # size-independent property at full size: total mass <= points, every value is k/5
  big = synth.make_point_clouds(8, seed=99).cuda()
  out = ops.pillar_scatter(big, use_ground_plane=True)
  vals = torch.unique(torch.round(out * 5))
  assert set(vals.tolist()) <= {0.0, 1.0, 2.0, 3.0, 4.0, 5.0}
  assert float(out.sum()) * 5 <= 8 * synth.N_POINTS


# ------------------------------------------------------------------------------------------------ tcgen05 GEMM
@pytest.mark.parametrize('rows,k,n', [(128, 64, 16), (256, 128, 64), (320, 72, 216), (1000, 216, 72), (640, 1512, 576),
                                      (352, 256, 2048), (64, 2048, 256), (37, 576, 1512)])
def test_linear_shapes(ops, rows, k, n):
  x, w, b = bf(rnd(rows, k, seed=1)), bf(rnd(n, k, seed=2, scale=k**-0.5)), rnd(n, seed=3)
  want = x.float() @ w.float().t() + b
  got = ops.linear(x, w, bias=b, out_f32=True)
  torch.cuda.synchronize()
  assert rel(got, want) < 2e-3
  got = ops.linear(x, w, bias=b, act=ops.ACT_RELU)
  assert rel(got.float(), F.relu(want)) < 6e-3


def test_linear_residual_rowmap_stats(ops):
  rows, k, n = 4 * 64, 216, 72
  x, w = bf(rnd(rows, k, seed=4)), bf(rnd(n, k, seed=5, scale=k**-0.5))
  res = rnd(4 * 320, n, seed=6)
  out = torch.zeros(4 * 320, n, device='cuda')
  pos = rnd(320, n, seed=7)
  ops.linear(x, w, res=res, out=out, row_map=(64, 320), res2=pos[256:], res2_strides=(0, 0, n, 1))
  want = torch.zeros_like(out)
  y = (x.float() @ w.float().t()).view(4, 64, n)
  want.view(4, 320, n)[:, :64] = y + res.view(4, 320, n)[:, :64] + pos[256:]
  assert rel(out, want) < 2e-3
  s, q = torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
  ops.linear(x, w, stats=(s, q), out_f32=True)
  yy = x.float() @ w.float().t()
  assert rel(s, yy.sum(0)) < 1e-3 and rel(q, (yy * yy).sum(0)) < 1e-3


def conv_ref(x_nhwc, w, stride=1, groups=1):
  return F.conv2d(x_nhwc.float().permute(0, 3, 1, 2), w.float(), None, stride=stride, padding=w.shape[-1] // 2,
                  groups=groups).permute(0, 2, 3, 1)


@pytest.mark.parametrize('b,h,w,cin,cout', [(2, 8, 32, 1512, 128), (2, 64, 64, 64, 320), (1, 16, 64, 32, 32),
                                            (3, 8, 8, 72, 64), (1, 32, 128, 128, 7)])
def test_conv3x3(ops, b, h, w, cin, cout):
  x = bf(rnd(b, h, w, cin, seed=8))
  wt = rnd(cout, cin, 3, 3, seed=9, scale=(9 * cin)**-0.5)
  bias = rnd(cout, seed=10)
  wp = ops.pack_conv_weight(wt)
  got = ops.conv_gemm(x, wp, taps=ops.TAPS_3X3, shift=bias, act=ops.ACT_RELU)
  want = F.relu(conv_ref(x, wp.float().view(cout, 3, 3, cin).permute(0, 3, 1, 2)) + bias)
  assert rel(got.float(), want) < 6e-3
  got = ops.conv_gemm(x, wp, taps=ops.TAPS_3X3, shift=bias, out_layout='nchw', out_f32=True)
  want = (conv_ref(x, wp.float().view(cout, 3, 3, cin).permute(0, 3, 1, 2)) + bias).permute(0, 3, 1, 2)
  assert rel(got, want) < 2e-3


@pytest.mark.parametrize('b,h,w,c', [(2, 16, 64, 72), (1, 32, 32, 216), (2, 8, 8, 1512), (1, 8, 32, 576)])
def test_grouped_conv(ops, b, h, w, c):
  x = bf(rnd(b, h, w, c, seed=11))
  wt = rnd(c, 24, 3, 3, seed=12, scale=(9 * 24)**-0.5)
  wp = ops.pack_grouped_conv_weight(wt)
  got = ops.conv_gemm(x, wp, taps=ops.TAPS_3X3, k_per_tile=48, a_c_per_ntile=48, bn=48)
  want = conv_ref(x, bf(wt), groups=c // 24)
  assert rel(got.float(), want) < 6e-3
  # stride 2 through parity planes
  xs = ops.parity_split(x)
  got = ops.conv_gemm(xs, wp, batch=b, taps=ops.taps_3x3_stride2(b), k_per_tile=48, a_c_per_ntile=48, bn=48)
  want = conv_ref(x, bf(wt), stride=2, groups=c // 24)
  assert got.shape == want.shape
  assert rel(got.float(), want) < 6e-3
  # 1x1 stride 2 (downsample) reads parity plane 0
  w1 = bf(rnd(40, c, seed=13, scale=c**-0.5))
  got = ops.conv_gemm(xs, w1.view(40, 1, c), batch=b)
  want = x.float()[:, ::2, ::2] @ w1.float().t()
  assert rel(got.float(), want) < 6e-3


# ------------------------------------------------------------------------------------------------ feature-map kernels
def test_stem_bn_se(ops):
  b = 2
  x = torch.randint(0, 256, (b, 3, 32, 64), generator=torch.Generator().manual_seed(1)).float().cuda()
  w = rnd(32, 3, 3, 3, seed=14, scale=0.2)
  a = torch.tensor([1 / (255 * 0.229), 1 / (255 * 0.224), 1 / (255 * 0.225)]).cuda()
  s = torch.tensor([-0.485 / 0.229, -0.456 / 0.224, -0.406 / 0.225]).cuda()
  st = (torch.zeros(32, device='cuda'), torch.zeros(32, device='cuda'))
  raw = ops.stem_conv(x, w, a, s, stats=st)
  xn = x * a.view(1, 3, 1, 1) + s.view(1, 3, 1, 1)
  want = F.conv2d(xn, w, None, stride=2, padding=1)
  assert rel(raw.float().permute(0, 3, 1, 2), want) < 4e-3
  assert rel(st[0], want.sum((0, 2, 3))) < 1e-3 and rel(st[1], (want * want).sum((0, 2, 3))) < 1e-3
  gamma, beta = rnd(32, seed=15).abs() + 0.5, rnd(32, seed=16)
  rm, rv = torch.zeros(32, device='cuda'), torch.ones(32, device='cuda')
  cnt = b * 16 * 32
  scale, shift, mean, invstd = ops.bn_finalize(st[0], st[1], gamma, beta, rm, rv, cnt, save=True)
  rm2, rv2 = torch.zeros(32, device='cuda'), torch.ones(32, device='cuda')
  want_bn = F.batch_norm(want, rm2, rv2, gamma, beta, training=True, momentum=0.1, eps=1e-5)
  assert rel(rm, rm2) < 1e-3 and rel(rv, rv2) < 1e-3
  pool = torch.zeros(b, 32, device='cuda')
  y = ops.scale_shift_act(raw, scale, shift, ops.ACT_RELU, pool_sum=pool)
  assert rel(y.float().permute(0, 3, 1, 2), F.relu(want_bn)) < 1e-2
  assert rel(pool / (16 * 32), y.float().mean((1, 2))) < 1e-3
  w1, b1, w2, b2 = rnd(8, 32, seed=17), rnd(8, seed=18), rnd(32, 8, seed=19), rnd(32, seed=20)
  gate = ops.se_gate(pool, 16 * 32, w1, b1, w2, b2)
  m = y.float().mean((1, 2))
  want_gate = torch.sigmoid(F.relu(m @ w1.t() + b1) @ w2.t() + b2)
  assert rel(gate, want_gate) < 1e-4
  z = ops.channel_scale(y, gate)
  assert rel(z.float(), y.float() * want_gate[:, None, None, :]) < 5e-3
  # eval-mode fused path
  y2 = ops.stem_conv(x, w, a, s, scale=scale, shift=shift, act=ops.ACT_RELU)
  assert rel(y2.float(), y.float()) < 1e-2


def test_pool_bilinear_layout(ops):
  b, h, w, c = 2, 16, 64, 72
  x = bf(rnd(b, h, w, c, seed=21))
  pos = rnd(320, c, seed=22)
  tok = torch.zeros(b, 320, c, device='cuda')
  ops.avgpool_tokens(x, tok, 8, 32, 0, pos_emb=pos)
  want = F.adaptive_avg_pool2d(x.float().permute(0, 3, 1, 2), (8, 32)).permute(0, 2, 3, 1).reshape(b, 256, c) + pos[:256]
  assert rel(tok[:, :256], want) < 1e-5
  up = ops.bilinear(tok, b, 8, 32, h, w, c, src_batch_stride=320 * c, src_row_stride=c, add=x)
  want = x.float() + F.interpolate(tok[:, :256].view(b, 8, 32, c).permute(0, 3, 1, 2), size=(h, w), mode='bilinear',
                                   align_corners=False).permute(0, 2, 3, 1)
  assert rel(up.float(), want) < 5e-3
  up2 = ops.bilinear(x, b, h, w, h * 4, w * 4, c)
  want = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=4, mode='bilinear',
                       align_corners=False).permute(0, 2, 3, 1)
  assert rel(up2.float(), want) < 5e-3
  mask = (rnd(64, 256, seed=23) > 0).float()
  o = ops.bilinear_nchw_mask(x, 11, 64, 256, mask)
  want = F.interpolate(x.float().permute(0, 3, 1, 2)[:, :11], size=(64, 256), mode='bilinear',
                       align_corners=False) * mask
  assert rel(o, want) < 1e-5
  xn = rnd(b, c, h, w, seed=24)
  assert rel(ops.nhwc_to_nchw(ops.nchw_to_nhwc(xn)), bf(xn).float()) == 0.0


def test_layernorm_attention(ops):
  rows, c = 640, 216
  x = rnd(rows, c, seed=25)
  g, bt = rnd(c, seed=26).abs() + 0.5, rnd(c, seed=27)
  yb, yf, _, _ = ops.layernorm(x, g, bt, want_f32=True)
  want = F.layer_norm(x, (c,), g, bt, 1e-5)
  assert rel(yf, want) < 1e-5 and rel(yb.float(), want) < 5e-3
  for c, b, t in ((72, 2, 320), (216, 2, 320), (576, 2, 320), (1512, 2, 320), (72, 3, 64), (216, 1, 160)):
    heads = 4
    qkv = bf(rnd(b, t, 3 * c, seed=28))
    out = ops.fusion_attn(qkv, b, t, c, heads).view(b, t, c)
    q, k, v = [u.float().view(b, t, heads, c // heads).transpose(1, 2) for u in qkv.split(c, dim=2)]
    att = F.softmax(q @ k.transpose(-2, -1) / math.sqrt(c // heads), dim=-1)
    want = (att @ v).transpose(1, 2).reshape(b, t, c)
    assert rel(out.float(), want) < 1e-2, c


def test_planner_kernels(ops):
  b, d = 3, 256
  q, mem = bf(rnd(b, 11, d, seed=29)), bf(rnd(b, 65, 2 * d, seed=30))
  out = ops.small_mha(q, mem, mem, b, 8, 11, 65, 32, (11 * d, d), (65 * 2 * d, 2 * d), (65 * 2 * d, 2 * d), v_off=d)
  qh = q.float().view(b, 11, 8, 32).transpose(1, 2)
  kh = mem.float()[..., :d].reshape(b, 65, 8, 32).transpose(1, 2)
  vh = mem.float()[..., d:].reshape(b, 65, 8, 32).transpose(1, 2)
  want = (F.softmax(qh @ kh.transpose(-2, -1) / math.sqrt(32), -1) @ vh).transpose(1, 2).reshape(b * 11, d)
  assert rel(out.float(), want) < 5e-3
  # extra sensor token (eval and batch-stat modes)
  vel, cmd = rnd(b, 1, seed=31).abs() * 4, F.one_hot(torch.tensor([0, 3, 5]), 6).float().cuda()
  w0, b0, w1, b1, pos = rnd(128, 7, seed=32), rnd(128, seed=33), rnd(256, 128, seed=34, scale=0.1), rnd(256, seed=35), \
      rnd(1, 256, seed=36)
  mem_f = torch.zeros(b, 65, 256, device='cuda')
  ops.extra_sensor_token(vel, cmd, 2.0, 1.5, False, None, None, w0, b0, w1, b1, pos, None, mem_f, 65, 64)
  vn = (vel - 2.0) / math.sqrt(1.5 + 1e-5)
  want = F.relu(F.relu(torch.cat([vn, cmd], 1) @ w0.t() + b0) @ w1.t() + b1) + pos
  assert rel(mem_f[:, 64], want) < 1e-5
  ops.extra_sensor_token(vel, cmd, 0.0, 1.0, True, None, None, w0, b0, w1, b1, pos, None, mem_f, 65, 64)
  vn = (vel - vel.mean()) / torch.sqrt(vel.var(unbiased=False) + 1e-5)
  want = F.relu(F.relu(torch.cat([vn, cmd], 1) @ w0.t() + b0) @ w1.t() + b1) + pos
  assert rel(mem_f[:, 64], want) < 1e-5
  # GRU + target speed vs torch.nn.GRU, evaluated on the CPU in float64 (cuDNN's GRU would run TF32 matmuls)
  torch.manual_seed(1234)
  gru = torch.nn.GRU(256, 64, batch_first=True).double()
  enc, dec = torch.nn.Linear(2, 64).double(), torch.nn.Linear(64, 2).double()
  ts = torch.nn.Sequential(torch.nn.Linear(256, 256), torch.nn.ReLU(), torch.nn.Linear(256, 4)).double()
  joined, tp = rnd(b, 11, 256, seed=37), rnd(b, 2, seed=38) * 10
  with torch.no_grad():
    jc, tc = joined.double().cpu(), tp.double().cpu()
    o, _ = gru(jc[:, :10], enc(tc).unsqueeze(0))
    want_cp = torch.cumsum(dec(o), 1)
    want_ts = ts(jc[:, 10])
  f = lambda t: t.detach().float().cuda()
  cp, tsl = ops.planner_head(joined, tp, f(enc.weight), f(enc.bias), f(gru.weight_ih_l0), f(gru.weight_hh_l0),
                             f(gru.bias_ih_l0), f(gru.bias_hh_l0), f(dec.weight), f(dec.bias), f(ts[0].weight),
                             f(ts[0].bias), f(ts[2].weight), f(ts[2].bias))
  assert rel(cp, want_cp) < 1e-4 and rel(tsl, want_ts) < 1e-4


def test_decode_heatmap(ops):
  from oracle import tfpp_oracle as orc
  b = 3
  maps = torch.rand(b, 21, 64, 64, generator=torch.Generator().manual_seed(3)).cuda()
  maps[:, 20] -= 0.5
  heat, wh, off, ycls, yres = maps[:, 0:4], maps[:, 4:6], maps[:, 6:8], maps[:, 8:20], maps[:, 20:21]
  got = ops.decode_heatmap(heat, wh, off, ycls, yres).cpu()
  want = orc.decode_heatmap(heat.cpu().contiguous(), wh.cpu().contiguous(), off.cpu().contiguous(),
                            ycls.cpu().contiguous(), yres.cpu().contiguous())
  assert got.shape == (b, 100, 9)
  assert torch.equal(got[..., 8], want[..., 8])          # scores bit-exact
  assert torch.equal(got[..., 7], want[..., 7])          # classes
  assert rel(got, want) < 1e-6
  g = np.load(os.path.join(GOLDEN, 'forward_eval_b2.npz'))
  assert g['boxes'].shape == (2, 100, 9)


# ------------------------------------------------------------------------------------------------ backward GEMMs
@pytest.mark.parametrize('b,h,w,cin,cout', [(2, 8, 32, 1512, 128), (2, 64, 64, 64, 320), (4, 16, 64, 72, 216),
                                            (3, 8, 8, 576, 576), (1, 1, 640, 216, 864)])
def test_wgrad_dense(ops, b, h, w, cin, cout):
  x, dy = bf(rnd(b, h, w, cin, seed=40)), bf(rnd(b, h, w, cout, seed=41))
  xf = x.float().permute(0, 3, 1, 2).requires_grad_(True)
  for k, taps in ((1, ops.TAPS_1X1), (3, ops.TAPS_3X3)):
    if k == 3 and h == 1:
      continue
    wt = torch.zeros(cout, cin, k, k, device='cuda', requires_grad=True)
    y = F.conv2d(xf, wt, None, padding=k // 2)
    (gw,) = torch.autograd.grad(y, wt, dy.float().permute(0, 3, 1, 2))
    got = ops.conv_wgrad(dy, x, taps=taps)  # (cout, taps, cin)
    want = gw.permute(0, 2, 3, 1).reshape(cout, k * k, cin)
    assert rel(got, want) < 2e-3, (k, rel(got, want))


@pytest.mark.parametrize('b,h,w,c', [(2, 16, 64, 72), (1, 32, 32, 216), (2, 8, 8, 1512)])
def test_wgrad_grouped_and_stride2(ops, b, h, w, c):
  x, dy = bf(rnd(b, h, w, c, seed=42)), bf(rnd(b, h, w, c, seed=43))
  xf = x.float().permute(0, 3, 1, 2)
  wt = torch.zeros(c, 24, 3, 3, device='cuda', requires_grad=True)
  y = F.conv2d(xf, wt, None, padding=1, groups=c // 24)
  (gw,) = torch.autograd.grad(y, wt, dy.float().permute(0, 3, 1, 2))
  got = ops.conv_wgrad(dy, x, taps=ops.TAPS_3X3, group_width=24)
  assert rel(got, gw.permute(0, 2, 3, 1).reshape(c, 9, 24)) < 2e-3
  # stride 2: dy at half resolution, x as parity planes
  dy2 = bf(rnd(b, h // 2, w // 2, c, seed=44))
  y = F.conv2d(xf, wt, None, stride=2, padding=1, groups=c // 24)
  (gw,) = torch.autograd.grad(y, wt, dy2.float().permute(0, 3, 1, 2))
  got = ops.conv_wgrad(dy2, ops.parity_split(x), taps=ops.taps_3x3_stride2(b), group_width=24)
  assert rel(got, gw.permute(0, 2, 3, 1).reshape(c, 9, 24)) < 2e-3


# ------------------------------------------------------------------------------------------------ small-channel convs
@pytest.mark.parametrize('cin,cout,h,w', [(32, 32, 64, 256), (32, 7, 48, 96), (32, 1, 40, 70), (16, 32, 64, 128)])
def test_smallc_conv3x3(ops, cin, cout, h, w):
  b = 2
  x = bf(rnd(b, h, w, cin, seed=50))
  wt = rnd(cout, cin, 3, 3, seed=51, scale=(9 * cin)**-0.5)
  bias = rnd(cout, seed=52)
  cpad = 8 if cout <= 8 else (16 if cout <= 16 else 32)
  wp = torch.nn.functional.pad(ops.pack_conv_weight(wt), (0, 0, 0, 0, 0, cpad - cout)).contiguous()
  want = conv_ref(x, bf(wt)) + bias  # NHWC
  if cout == cpad:
    got = ops.smallc_conv3x3(x, wp, bias=bias, act=ops.ACT_RELU)
    assert rel(got.float(), F.relu(want)) < 6e-3
  got = ops.smallc_conv3x3(x, wp, bias=bias, n_valid=cout, out_nchw_f32=True, act=ops.ACT_SIGMOID, act_n_limit=1)
  w2 = want.permute(0, 3, 1, 2).clone()
  w2[:, :1] = torch.sigmoid(w2[:, :1])
  assert got.shape == (b, cout, h, w)
  assert rel(got, w2) < 2e-3


@pytest.mark.parametrize('cop,co_valid,h,w', [(32, 32, 64, 256), (16, 7, 48, 96), (16, 1, 40, 70)])
def test_smallc_wgrad_and_dgrad(ops, cop, co_valid, h, w):
  b, cin = 2, 32
  x = bf(rnd(b, h, w, cin, seed=53))
  dy = bf(rnd(b, h, w, cop, seed=54))
  dy[..., co_valid:] = 0
  wt = torch.zeros(co_valid, cin, 3, 3, device='cuda', requires_grad=True)
  xin = x.float().permute(0, 3, 1, 2).requires_grad_(True)
  y = F.conv2d(xin, wt, None, padding=1)
  gw, = torch.autograd.grad(y, wt, dy.float()[..., :co_valid].permute(0, 3, 1, 2))
  out = torch.zeros(co_valid, cin, 3, 3, device='cuda')
  ops.smallc_wgrad3x3(dy, x, out, (cin * 9, 1, 9), co_valid)
  assert rel(out, gw) < 2e-3
  # dgrad through the same forward kernel with the flipped / transposed pack
  wr = rnd(co_valid, cin, 3, 3, seed=55, scale=0.1)
  wd = wr.flip(2, 3).permute(1, 2, 3, 0).reshape(cin, 9, co_valid)
  wd = torch.nn.functional.pad(wd, (0, cop - co_valid)).to(torch.bfloat16).contiguous()
  got = ops.smallc_conv3x3(dy, wd)
  y2 = F.conv2d(xin, bf(wr).float(), None, padding=1)
  gx, = torch.autograd.grad(y2, xin, dy.float()[..., :co_valid].permute(0, 3, 1, 2))
  assert rel(got.float().permute(0, 3, 1, 2), gx) < 6e-3


# ------------------------------------------------------------------------------------------------ feature-map adjoints
@pytest.mark.parametrize('b,h,w,c,act,with_se', [(3, 9, 13, 72, 1, True), (2, 8, 8, 1512, 1, False), (2, 17, 5, 216, 0, False),
                                                (5, 4, 4, 576, 1, True)])
def test_bn_backward(ops, b, h, w, c, act, with_se):
  """tfpp_bn_bwd (two passes) against autograd of batch_norm(+relu)(*gate, +pool) on the same bf16 tensors."""
  raw = bf(rnd(b, h, w, c, seed=1, scale=2.0) + 0.5)
  gamma, beta = rnd(c, seed=2).abs() + 0.5, rnd(c, seed=3) * 0.1
  dy = bf(rnd(b, h, w, c, seed=4))
  gate = torch.sigmoid(rnd(b, c, seed=5)) if with_se else None
  pgrad = rnd(b, c, seed=6) * 0.05 if with_se else None
  rawf = raw.float().requires_grad_(True)
  g32, b32 = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
  mean = rawf.detach().mean((0, 1, 2))
  var = rawf.detach().var((0, 1, 2), unbiased=False)
  invstd = torch.rsqrt(var + 1e-5)
  z = F.batch_norm(rawf.permute(0, 3, 1, 2), None, None, g32, b32, training=True, eps=1e-5).permute(0, 2, 3, 1)
  y = F.relu(z) if act else z
  yb = bf(y.detach())
  loss = (y * (gate[:, None, None, :] if with_se else 1.0) * dy.float()).sum()
  if with_se:
    loss = loss + (y.sum((1, 2)) * pgrad).sum()
  loss.backward()
  dgamma, dbeta = torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
  draw, dz = ops.bn_bwd(dy, yb if act else None, raw, mean, invstd, gamma, act, dgamma, dbeta, gate=gate, pool_grad=pgrad,
                        want_dz=True)
  assert rel(draw.float(), rawf.grad) < 1e-2
  assert rel(dgamma, g32.grad) < 2e-3 and rel(dbeta, b32.grad) < 2e-3
  want_dz = dy.float() * (gate[:, None, None, :] if with_se else 1.0) + (pgrad[:, None, None, :] if with_se else 0.0)
  if act:
    want_dz = want_dz * (yb.float() > 0)
  assert rel(dz.float(), want_dz) < 5e-3
  if act:  # same result with the ReLU mask recomputed from raw and the forward affine instead of read from y
    scale = gamma * invstd
    shift = beta - mean * scale
    yb2 = ops.scale_shift_act(raw, scale, shift, ops.ACT_RELU)
    dg2, db2 = torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
    dg1, db1 = torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
    d1, _ = ops.bn_bwd(dy, yb2, raw, mean, invstd, gamma, act, dg1, db1, gate=gate, pool_grad=pgrad)
    d2, _ = ops.bn_bwd(dy, None, raw, mean, invstd, gamma, act, dg2, db2, gate=gate, pool_grad=pgrad,
                       fwd_affine=(scale, shift))
    assert rel(d2.float(), d1.float()) < 1e-3 and rel(dg2, dg1) < 1e-4 and rel(db2, db1) < 1e-4


@pytest.mark.parametrize('b,c,rd,hw', [(5, 1512, 144, 64), (2, 72, 8, 35), (9, 216, 18, 16), (32, 576, 54, 4)])
def test_se_forward_backward(ops, b, c, rd, hw):
  """SE excite MLP (two contraction kernels each way) against autograd."""
  h = w = None
  for hh in range(1, hw + 1):
    if hw % hh == 0:
      h, w = hh, hw // hh
  a2 = bf(rnd(b, h, w, c, seed=1).abs())
  dout = bf(rnd(b, h, w, c, seed=2))
  w1, b1 = rnd(rd, c, seed=3, scale=c ** -0.5), rnd(rd, seed=4, scale=0.1)
  w2, b2 = rnd(c, rd, seed=5, scale=rd ** -0.5), rnd(c, seed=6, scale=0.1)
  pool = a2.float().sum((1, 2)).requires_grad_(True)
  p = [t.clone().requires_grad_(True) for t in (w1, b1, w2, b2)]
  hid = F.relu((pool / hw) @ p[0].t() + p[1])
  gate_want = torch.sigmoid(hid @ p[2].t() + p[3])
  gate, hidden = ops.se_gate(pool.detach(), hw, w1, b1, w2, b2, want_hidden=True)
  assert rel(gate, gate_want) < 1e-5 and rel(hidden, hid) < 1e-5
  (a2.float() * gate_want[:, None, None, :] * dout.float()).sum().backward()
  dws = [torch.zeros_like(t) for t in (w1, b1, w2, b2)]
  pool_grad = ops.se_bwd(dout, a2, gate, hidden, pool.detach(), hw, w1, w2, *dws)
  assert rel(pool_grad, pool.grad) < 1e-4
  for got, want in zip(dws, p):
    assert rel(got, want.grad) < 1e-4


@pytest.mark.parametrize('rows,c,layout,act,limit', [(77, 4536, 0, 0, 0), (300, 64, 0, 1, 0), (41, 1512, 2, 0, 0),
                                                    (130, 24, 0, 2, 16), (64, 20, 0, 1, 0), (50, 576, 2, 1, 0)])
def test_act_backward(ops, rows, c, layout, act, limit):
  """tfpp_act_bwd: vector path (C % 8 == 0) and scalar fallback, bf16 / fp32 incoming gradients, bias reduction."""
  dy32 = rnd(rows, c, seed=1)
  dy = dy32 if layout == 2 else bf(dy32)
  y = bf(rnd(rows, c, seed=2)) if act == 1 else bf(torch.sigmoid(rnd(rows, c, seed=2)))
  dbias = torch.zeros(c, device='cuda')
  dz = ops.act_bwd(dy, y if act else None, act, 1, rows, c, layout=layout, act_n_limit=limit, dy_scale=0.5, dbias=dbias)
  want = dy.float() * 0.5
  if act == 1:
    want = want * (y.float() > 0)
  elif act == 2:
    m = torch.ones(c, device='cuda') if limit == 0 else (torch.arange(c, device='cuda') < limit).float()
    yf = y.float()
    want = want * (m * yf * (1 - yf) + (1 - m))
  assert rel(dz.float(), want) < 4e-3
  assert rel(dbias, want.sum(0)) < 1e-4


# ------------------------------------------------------------------------------------------------ haloed-tile group conv
@pytest.mark.parametrize('b,h,w,c,stride', [(2, 16, 64, 72, 1), (1, 8, 8, 216, 1), (3, 20, 12, 144, 1), (2, 64, 64, 72, 2),
                                            (1, 16, 16, 576, 2), (2, 10, 36, 72, 2), (1, 8, 32, 1512, 1)])
def test_gconv3x3_forward_stats_affine(ops, b, h, w, c, stride):
  """tfpp_gconv3x3 against F.conv2d(groups=C/24) on the bf16-rounded operands; BatchNorm statistics; eval affine+ReLU."""
  x = bf(rnd(b, h, w, c, seed=1))
  wt = rnd(c, 24, 3, 3, seed=2, scale=0.1)
  wp = ops.pack_gconv_halo(wt)
  want = F.conv2d(x.float().permute(0, 3, 1, 2), bf(wt).float(), stride=stride, padding=1, groups=c // 24)
  st = (torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda'))
  y = ops.gconv3x3(x, wp, stride, stats=st)
  assert y.shape == (b, h // stride, w // stride, c)
  assert rel(y.float().permute(0, 3, 1, 2), want) < 4e-3
  assert rel(st[0], want.sum((0, 2, 3))) < 2e-3 and rel(st[1], (want * want).sum((0, 2, 3))) < 2e-3
  sc, sh = rnd(c, seed=3).abs() + 0.5, rnd(c, seed=4)
  y2 = ops.gconv3x3(x, wp, stride, scale=sc, shift=sh, act=ops.ACT_RELU)
  assert rel(y2.float().permute(0, 3, 1, 2), F.relu(want * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))) < 4e-3


@pytest.mark.parametrize('b,h,w,c', [(2, 16, 32, 72), (1, 8, 8, 216), (2, 24, 20, 144)])
def test_gconv3x3_input_gradient(ops, b, h, w, c):
  """stride-1 input gradient = the same kernel on dY with the transposed / flipped pack, against autograd."""
  x = rnd(b, c, h, w, seed=1).requires_grad_(True)
  wt = rnd(c, 24, 3, 3, seed=2, scale=0.1)
  dy = bf(rnd(b, h, w, c, seed=3))
  F.conv2d(x, bf(wt).float(), padding=1, groups=c // 24).backward(dy.float().permute(0, 3, 1, 2))
  got = ops.gconv3x3(dy, ops.pack_gconv_halo(wt, transpose=True))
  assert rel(got.float().permute(0, 3, 1, 2), x.grad) < 4e-3


@pytest.mark.parametrize('b,h,w,c,stride', [(2, 16, 32, 72, 1), (3, 8, 8, 216, 1), (2, 24, 20, 144, 1), (2, 32, 64, 72, 2),
                                            (1, 16, 16, 216, 2), (2, 12, 40, 72, 2)])
def test_gconv3x3_weight_gradient(ops, b, h, w, c, stride):
  """tfpp_gconv3x3_wgrad (+= into the torch-layout gradient) against autograd of F.conv2d(groups=C/24)."""
  x = bf(rnd(b, h, w, c, seed=1))
  dy = bf(rnd(b, h // stride, w // stride, c, seed=2))
  wt = rnd(c, 24, 3, 3, seed=3, scale=0.1).requires_grad_(True)
  F.conv2d(x.float().permute(0, 3, 1, 2), wt, stride=stride, padding=1, groups=c // 24).backward(
      dy.float().permute(0, 3, 1, 2))
  dw = torch.full((c, 24, 3, 3), 0.5, device='cuda')
  ops.gconv3x3_wgrad(dy, x, dw, stride)
  assert rel(dw - 0.5, wt.grad) < 3e-3


@pytest.mark.parametrize('c,b,t', [(72, 2, 320), (216, 2, 320), (576, 1, 320), (1512, 1, 320), (72, 3, 64), (216, 1, 160)])
def test_fusion_attention_backward(ops, c, b, t):
  """tfpp_fusion_attn_bwd against autograd of softmax(QK^T/sqrt(hd))V on the same bf16 inputs."""
  heads = 4
  qkv = bf(rnd(b, t, 3 * c, seed=40))
  dout = bf(rnd(b, t, c, seed=41))
  x = qkv.float().requires_grad_(True)
  q, k, v = [u.view(b, t, heads, c // heads).transpose(1, 2) for u in x.split(c, dim=2)]
  att = F.softmax(q @ k.transpose(-2, -1) / math.sqrt(c // heads), dim=-1)
  (att @ v).transpose(1, 2).reshape(b, t, c).backward(dout.float())
  got = ops.fusion_attn_bwd(qkv, dout, b, t, c, heads).view(b, t, 3 * c).float()
  for name, sl in (('dq', slice(0, c)), ('dk', slice(c, 2 * c)), ('dv', slice(2 * c, 3 * c))):
    assert rel(got[..., sl], x.grad[..., sl]) < 2e-2, (name, c)


@pytest.mark.parametrize('b,ho,wo,c', [(2, 16, 32, 72), (1, 8, 8, 216), (2, 12, 20, 144), (1, 4, 16, 576)])
def test_gconv3x3_stride2_input_gradient(ops, b, ho, wo, c):
  """tfpp_gconv3x3_dgrad_s2 (parity classes on a dY tile) against autograd of the stride-2 group conv."""
  x = rnd(b, c, 2 * ho, 2 * wo, seed=1).requires_grad_(True)
  wt = rnd(c, 24, 3, 3, seed=2, scale=0.1)
  dy = bf(rnd(b, ho, wo, c, seed=3))
  F.conv2d(x, bf(wt).float(), stride=2, padding=1, groups=c // 24).backward(dy.float().permute(0, 3, 1, 2))
  got = ops.gconv3x3_dgrad_s2(dy, ops.pack_gconv_halo(wt, transpose=True))
  assert got.shape == (b, 2 * ho, 2 * wo, c)
  assert rel(got.float().permute(0, 3, 1, 2), x.grad) < 4e-3


# ------------------------------------------------------------------------------------------------ tcgen05 haloed-tile conv
# (op-level parity green on B200, profiles/r01_halo_umma_optest_v19.log; the engine only uses it with TFPP_HALO_UMMA=1)
@pytest.mark.parametrize('cin,cout,n_pad,h,w', [(32, 32, 32, 16, 128), (32, 7, 16, 24, 70), (16, 32, 32, 8, 62),
                                                (64, 32, 32, 9, 130)])
def test_halo_umma_conv3x3(ops, cin, cout, n_pad, h, w):
  """tfpp_halo_conv3x3 (tcgen05 over channel-chunk-major halo planes) against F.conv2d; forward and the dgrad pack."""
  b = 2
  x = bf(rnd(b, h, w, cin, seed=1))
  wt = rnd(cout, cin, 3, 3, seed=2, scale=0.1)
  bias = rnd(cout, seed=3)
  want = F.conv2d(x.float().permute(0, 3, 1, 2), bf(wt).float(), bias, padding=1)
  y = ops.halo_conv3x3(x, ops.pack_halo_umma_weight(wt, n_pad), bias=bias, n_valid=cout)
  assert rel(y.float()[..., :cout].permute(0, 3, 1, 2), want) < 4e-3
  y2 = ops.halo_conv3x3(x, ops.pack_halo_umma_weight(wt, n_pad), bias=bias, n_valid=cout, act=ops.ACT_RELU,
                        out_nchw_f32=True)
  assert rel(y2, F.relu(want)) < 4e-3
  if cout % 8 == 0 and cout in (16, 32, 64):
    dy = bf(rnd(b, h, w, cout, seed=4))
    xg = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.conv2d(xg, bf(wt).float(), padding=1).backward(dy.float().permute(0, 3, 1, 2))
    cin_pad = cin
    dx = ops.halo_conv3x3(dy, ops.pack_halo_umma_weight(wt, cin_pad, transpose=True))
    assert rel(dx.float().permute(0, 3, 1, 2), xg.grad) < 4e-3


@pytest.mark.skipif(os.environ.get('TFPP_EXPERIMENTAL', '0') != '1', reason='not yet run on a GPU: opt-in (round 2)')
@pytest.mark.parametrize('b,h,w,c', [(2, 16, 64, 72), (1, 8, 8, 216), (3, 20, 12, 144), (1, 8, 130, 1512)])
def test_halo_umma_gconv3x3(ops, b, h, w, c):
  """tfpp_halo_gconv3x3 (tcgen05 group conv over haloed planes) against F.conv2d(groups=C/24): raw output + BatchNorm
  statistics, eval affine + ReLU, and the stride-1 input gradient through the transposed pack."""
  x = bf(rnd(b, h, w, c, seed=1))
  wt = rnd(c, 24, 3, 3, seed=2, scale=0.1)
  want = F.conv2d(x.float().permute(0, 3, 1, 2), bf(wt).float(), padding=1, groups=c // 24)
  st = (torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda'))
  y = ops.halo_gconv3x3(x, ops.pack_halo_gconv_weight(wt), stats=st)
  assert rel(y.float().permute(0, 3, 1, 2), want) < 4e-3
  assert rel(st[0], want.sum((0, 2, 3))) < 2e-3 and rel(st[1], (want * want).sum((0, 2, 3))) < 2e-3
  sc, sh = rnd(c, seed=3).abs() + 0.5, rnd(c, seed=4)
  y2 = ops.halo_gconv3x3(x, ops.pack_halo_gconv_weight(wt), scale=sc, shift=sh, act=ops.ACT_RELU)
  assert rel(y2.float().permute(0, 3, 1, 2), F.relu(want * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))) < 4e-3
  xg = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
  dy = bf(rnd(b, h, w, c, seed=5))
  F.conv2d(xg, bf(wt).float(), padding=1, groups=c // 24).backward(dy.float().permute(0, 3, 1, 2))
  dx = ops.halo_gconv3x3(dy, ops.pack_halo_gconv_weight(wt, transpose=True))
  assert rel(dx.float().permute(0, 3, 1, 2), xg.grad) < 4e-3


def test_pillar_scatter_with_fused_alignment_bit_exact(ops):
  """tfpp_pillar_scatter_aligned == CARLA_Data.align + lidar_to_histogram_features of the unmodified reference
  (tests/golden/make_align_golden.py): the two rigid transforms are applied per point in float64 inside K1."""
  from carla_garage_b200 import dataio, synth
  g = np.load(os.path.join(GOLDEN, 'align.npz'))
  cases = g['cases']
  pts = synth.make_point_clouds(len(cases), seed=21, n_points=20000)
  xf = np.stack([dataio.align_transforms({'pos_global': (c[0], c[1]), 'theta': c[2]}, {'pos_global': (c[3], c[4]), 'theta': c[5]},
                                         y_augmentation=c[6], yaw_augmentation=c[7]) for c in cases])
  for gp in (0, 1):
    out = ops.pillar_scatter(pts.cuda(), use_ground_plane=bool(gp), xform=torch.from_numpy(xf).cuda()).cpu().numpy()
    for i in range(len(cases)):
      want = g[f'hist{i}_gp{gp}'].astype(np.float32) / 5.0
      diff = int((out[i] != want).sum())
      assert diff == 0, (i, gp, diff)
  # identity transform == the plain kernel except where f32 z == float32(0.2) would flip (none in this cloud)
  ident = torch.zeros(1, 1, 4, dtype=torch.float64, device='cuda')
  a = ops.pillar_scatter(pts[2:3].cuda(), use_ground_plane=True, xform=ident)
  b = ops.pillar_scatter(pts[2:3].cuda(), use_ground_plane=True)
  assert torch.equal(a, b)
