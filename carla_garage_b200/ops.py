"""Thin torch-tensor wrappers over the C ABI (include/tfpp.h).  torch is used for device memory and the current
stream only; every computation happens in libtfpp.so.  No fallback: a missing library or a CPU tensor raises."""
import ctypes
import os

import torch

from . import _lib
from ._lib import ConvGemmArgs, WgradArgs, check

BF16 = torch.bfloat16
F32 = torch.float32
ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_GELU = 0, 1, 2, 3

# Activation storage dtype of the engine: bf16 (production) or fp32 (parity mode, csrc/fp32_path.cu).  Every wrapper
# below dispatches on the dtype of the tensor it is handed; ACT_DTYPE only decides what the sources (stem conv, weight
# packs, explicit buffers of engine.py) produce.
ACT_DTYPE = [BF16]


def set_precision(mode):
  """'bf16' (default, the benchmarked tcgen05 path) or 'fp32' (north_star's 1e-3 parity mode: fp32 feature maps and
  CUDA-core fp32 contractions, forward only)."""
  if mode not in ('bf16', 'fp32'):
    raise ValueError(mode)
  ACT_DTYPE[0] = F32 if mode == 'fp32' else BF16


def act_dtype():
  return ACT_DTYPE[0]


class precision:
  """with ops.precision('fp32'): ..."""

  def __init__(self, mode):
    self.mode = mode

  def __enter__(self):
    self.prev = ACT_DTYPE[0]
    set_precision(self.mode)

  def __exit__(self, *exc):
    ACT_DTYPE[0] = self.prev


TAPS_1X1 = ((0, 0, 0, 0),)
TAPS_3X3 = tuple((kx - 1, ky - 1, 0, ky * 3 + kx) for ky in range(3) for kx in range(3))


def taps_3x3_stride2(batch):
  """3x3 / stride 2 / pad 1 on parity planes (tfpp_parity_split): input row 2*oy+ky-1 lives in plane (ky+1)&1 at
  plane row oy + (-1 if ky == 0 else 0); same for columns."""
  taps = []
  for ky in range(3):
    for kx in range(3):
      py, dy = (1, -1) if ky == 0 else ((0, 0) if ky == 1 else (1, 0))
      px, dx = (1, -1) if kx == 0 else ((0, 0) if kx == 1 else (1, 0))
      taps.append((dx, dy, (py * 2 + px) * batch, ky * 3 + kx))
  return tuple(taps)


_PROFILE = None


def profile_gemm_launches(fn):
  """Run fn() once with CUDA events (on the launching stream) around every tcgen05 GEMM launch; returns the summed
  device time, the algorithmic FLOPs and the achieved TFLOP/s of that kernel family."""
  global _PROFILE  # pylint: disable=global-statement
  _PROFILE = []
  try:
    fn()
    torch.cuda.synchronize()
    ms = sum(e0.elapsed_time(e1) for e0, e1, _, _ in _PROFILE)
    flop = sum(f for _, _, f, _ in _PROFILE)
    n = len(_PROFILE)
    dump = os.environ.get('TFPP_GEMM_DUMP')
    if dump:  # per-launch shapes and times, aggregated by shape (kernel-tuning aid)
      agg = {}
      for e0, e1, f, desc in _PROFILE:
        a = agg.setdefault(desc, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += e0.elapsed_time(e1)
        a[2] += f
      with open(dump, 'w', encoding='utf-8') as fh:
        for desc, (cnt, t, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
          fh.write(f'{t:9.3f} ms  n={cnt:4d}  avg {t / cnt * 1e3:8.1f} us  {f / t / 1e9 if t else 0:8.1f} TFLOP/s  {desc}\n')
  finally:
    _PROFILE = None
  return {'ms': ms, 'gflop': flop / 1e9, 'launches': n, 'tflops': flop / (ms * 1e-3) / 1e12 if ms > 0 else 0.0}


def _stream():
  return torch.cuda.current_stream().cuda_stream


def _p(t):
  return None if t is None else t.data_ptr()


def _dev(t, dtype=None):
  if not t.is_cuda:
    raise RuntimeError('carla_garage_b200 ops need CUDA tensors (there is no CPU fallback)')
  if dtype is not None and t.dtype != dtype:
    raise RuntimeError(f'expected {dtype}, got {t.dtype}')
  if not t.is_contiguous():
    raise RuntimeError('expected a contiguous tensor')
  return t


def pick_tile(height, width):
  """128-pixel tile (tw, th, nb) of one conv_gemm CTA."""
  if width >= 128:
    return 128, 1, 1
  tw = 1
  while tw * 2 <= width:
    tw *= 2
  if tw < width:  # non power-of-two narrow maps: next power of two, masked
    tw *= 2
  th = 1
  while th * 2 <= height and tw * th * 2 <= 128:
    th *= 2
  if th < height and tw * th * 2 <= 128:
    th *= 2
  return tw, th, 128 // (tw * th)


def pick_bn(n):
  """N tile (multiple of 16, <= 256): the largest tile whose padded width is within 4 % of the minimum (large tiles
  re-read the A operand fewer times; padding columns are wasted MMA work)."""
  cands = [(-(-n // bn) * bn, bn) for bn in range(16, 257, 16)]
  min_pad = min(p for p, _ in cands)
  return max(bn for p, bn in cands if p <= min_pad * 1.04)


def pick_bn_for(n, m_tiles, sms=148):
  """N tile for a persistent GEMM over ``m_tiles`` 128-pixel tiles on ``sms`` CTAs.  The kernels are bound by the operand
  feed (a k-block moves 128 rows of A + bn rows of B), so the time is ~ rounds * (128 + bn): problems of many rounds keep
  pick_bn's wide tile; problems of one to three rounds (LiDAR branch, decoder / planner GEMMs) take the tile width that
  fills the last round (e.g. n = 576 on 64 pixel tiles: bn = 144 -> 256 tiles, two full-ish rounds of narrower tiles,
  instead of 192 -> 1.3 rounds)."""
  wide = pick_bn(n)
  if m_tiles * -(-n // wide) >= 4 * sms:
    return wide
  best, best_cost = wide, None
  for bn in range(32, 257, 16):
    tiles = m_tiles * -(-n // bn)
    cost = -(-tiles // sms) * (128 + bn) * (1.0 + 0.25 * (-(-n // bn) * bn - n) / n)   # padding columns are wasted work
    if best_cost is None or cost < best_cost - 1e-9 or (abs(cost - best_cost) <= 1e-9 and bn > best):
      best, best_cost = bn, cost
  return best


def nhwc_strides(h, w, c):
  return (h * w * c, w * c, c, 1)


def nchw_strides(h, w, c):
  return (c * h * w, w, 1, h * w)


def taps_3x3_stride2_dgrad(py, px):
  """Taps of the input gradient of a 3x3 / stride 2 / pad 1 conv for the input pixels of parity (py, px), as an implicit
  GEMM over the OUTPUT-gradient map: input pixel (2r + py, 2c + px) is read by the kernel rows ky whose parity plane is py
  — ky = 1 for even rows (output row r), ky in {0, 2} for odd rows (output rows r + 1 and r) — same for columns.
  Returns ((dx, dy, 0, ky * 3 + kx), ...) for ops.conv_gemm over the (Cin, 9, Cout) weight pack."""
  kys = ((1, 0),) if py == 0 else ((0, 1), (2, 0))
  kxs = ((1, 0),) if px == 0 else ((0, 1), (2, 0))
  return tuple((ox, oy, 0, ky * 3 + kx) for ky, oy in kys for kx, ox in kxs)


def conv_gemm(a, w, *, a_shape=None, a_batch_stride=0, batch=None, taps=TAPS_1X1, k_per_tile=None, a_c_per_ntile=0, bn=None, out=None, out_f32=False,
              out_layout='nhwc', out_strides=None, res1=None, res1_strides=None, res2=None, res2_strides=None,
              scale=None, shift=None, act=ACT_NONE, act_n_limit=0, stats=None, no_output=False, drop=None):
  """out[pixel, n] = act(scale[n] * sum_{tap,c} a[pixel + tap, c] * w[n, tap, c] + shift[n] + res1 + res2).

  a: (Ba, H, W, C) bf16 NHWC; w: (N, taps, K) bf16.  Returns a new (B,H,W,N) bf16 / (B,N,H,W) f32 tensor unless
  ``out`` (+ ``out_strides`` in elements (sb, sy, sx, sn)) is given.  stats = (sum, sumsq) f32 (N,) accumulators.
  """
  f32 = w.dtype == F32  # fp32 parity mode: same contract on tfpp_conv_gemm_f32
  _dev(w, F32 if f32 else BF16)
  if a.dtype != w.dtype:
    raise RuntimeError(f'conv_gemm: activations are {a.dtype}, weight pack is {w.dtype}')
  if a_shape is None:  # contiguous NHWC tensor
    _dev(a)
    ab, h, wd, c = a.shape
  else:  # strided slab inside a larger bf16 buffer: a_shape = (Ba, H, W, C), images a_batch_stride elements apart
    ab, h, wd, c = a_shape
  n, wt, kd = w.shape
  b = ab if batch is None else batch
  tw, th, nb = pick_tile(h, wd)
  args = ConvGemmArgs()
  args.a, args.a_batch, args.height, args.width, args.a_channels = a.data_ptr(), ab, h, wd, c
  args.a_batch_stride = a_batch_stride
  args.w, args.w_taps, args.w_kdim, args.n = w.data_ptr(), wt, kd, n
  args.batch = b
  args.k_per_tile = kd if k_per_tile is None else k_per_tile
  args.a_c_per_ntile = a_c_per_ntile
  if bn is None:
    bn = pick_bn_for(n, -(-wd // tw) * -(-h // th) * -(-b // nb))
  args.bn = bn
  args.tw, args.th, args.nb = tw, th, nb
  args.ntaps = len(taps)
  for i, (dx, dy, db, tw_) in enumerate(taps):
    args.tap_dx[i], args.tap_dy[i], args.tap_db[i], args.tap_w[i] = dx, dy, db, tw_
  if no_output:
    out = None
    args.out = None
  else:
    if out is None:
      if out_layout == 'nhwc':
        out = torch.empty((b, h, wd, n), dtype=F32 if (out_f32 or f32) else BF16, device=a.device)
        out_strides = nhwc_strides(h, wd, n)
      else:
        out = torch.empty((b, n, h, wd), dtype=F32 if (out_f32 or f32) else BF16, device=a.device)
        out_strides = nchw_strides(h, wd, n)
    args.out = out.data_ptr()
    args.out_f32 = int(out.dtype == F32)
    args.o_sb, args.o_sy, args.o_sx, args.o_sn = out_strides
  for name, r, rs in (('1', res1, res1_strides), ('2', res2, res2_strides)):
    if r is not None:
      if rs is None:
        rs = nhwc_strides(h, wd, n)
      setattr(args, f'res{name}', r.data_ptr())
      setattr(args, f'res{name}_f32', int(r.dtype == F32))
      for f, v in zip(('sb', 'sy', 'sx', 'sn'), rs):
        setattr(args, f'r{name}_{f}', v)
  args.scale = _p(scale)
  args.shift = _p(shift)
  args.act = act
  args.act_n_limit = act_n_limit
  if stats is not None:
    args.stat_sum, args.stat_sq = stats[0].data_ptr(), stats[1].data_ptr()
  if drop is not None:  # (rng tensor {seed, step}, p, site): out = drop(act(.)) + residuals
    args.drop_rng, args.drop_p, args.drop_site = drop[0].data_ptr(), drop[1], drop[2]
  if _PROFILE is not None:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
  if f32:
    check(_lib.load().tfpp_conv_gemm_f32(ctypes.byref(args), _stream()), 'tfpp_conv_gemm_f32')
  else:
    check(_lib.load().tfpp_conv_gemm(ctypes.byref(args), _stream()), 'tfpp_conv_gemm')
  if _PROFILE is not None:
    e1.record()
    k_alg = 24 if a_c_per_ntile else args.k_per_tile
    _PROFILE.append((e0, e1, 2.0 * b * h * wd * n * k_alg * len(taps),
                     f'gemm  b{b} {h}x{wd} n{n} k{kd} taps{len(taps)} bn{bn} grp{int(bool(a_c_per_ntile))} stats{int(stats is not None)}'))
  return out


def pick_tile64(height, width):
  tw, th, nb = pick_tile(height, width)
  if nb >= 2:
    return tw, th, nb // 2
  if th >= 2:
    return tw, th // 2, nb
  return tw // 2, th, nb


def conv_wgrad(dy, x, *, cin=None, taps=TAPS_1X1, w_taps=None, group_width=0, x_batch_stride=0, x_shape=None, bn=0,
               splits=0, out=None, out_strides=None, dy_shape=None, cout_valid=0):
  """dw[co, tap, ci] = sum_pixels dy[pixel, co] * x[pixel + tap, ci]; dy (B,H,W,Cout) bf16, x (Bx,H,W,Cx) bf16.
  Returns fp32 (Cout, w_taps, cin) (dense) or (Cout, w_taps, group_width) (grouped); accumulates into ``out``."""
  f32 = dy.dtype == F32   # fp32 parity mode
  if x.dtype != dy.dtype:
    raise RuntimeError(f'conv_wgrad: dy is {dy.dtype}, x is {x.dtype}')
  if dy_shape is None:
    _dev(dy)
    b, h, w, cout = dy.shape
  else:
    b, h, w, cout = dy_shape
  if x_shape is None:
    _dev(x)
    xb, _, _, cx = x.shape
  else:
    xb, _, _, cx = x_shape
  cin = cx if cin is None else cin
  w_taps = len(taps) if w_taps is None else w_taps
  if out is None:
    kdim = group_width if group_width else cin
    out = torch.zeros((cout_valid or cout, w_taps, kdim), dtype=F32, device=dy.device)
    out_strides = (w_taps * kdim, kdim, 1)
  a = WgradArgs()
  a.dy, a.x, a.dw = dy.data_ptr(), x.data_ptr(), out.data_ptr()
  a.batch, a.height, a.width, a.cout, a.cout_valid = b, h, w, cout, cout_valid
  a.x_batch, a.x_channels, a.x_batch_stride = xb, cx, x_batch_stride
  a.cin, a.group_width, a.ntaps = cin, group_width, len(taps)
  a.dw_s_co, a.dw_s_tap, a.dw_s_ci = out_strides
  for i, (dx, dy_, db, tw_) in enumerate(taps):
    a.tap_dx[i], a.tap_dy[i], a.tap_db[i], a.tap_w[i] = dx, dy_, db, tw_
  a.tw, a.th, a.nb = pick_tile64(h, w)
  a.bn, a.splits = bn, splits
  if _PROFILE is not None:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
  if f32:
    check(_lib.load().tfpp_conv_wgrad_f32(ctypes.byref(a), _stream()), 'tfpp_conv_wgrad_f32')
  else:
    check(_lib.load().tfpp_conv_wgrad(ctypes.byref(a), _stream()), 'tfpp_conv_wgrad')
  if _PROFILE is not None:
    e1.record()
    _PROFILE.append((e0, e1, 2.0 * b * h * w * (cout_valid or cout) * (group_width or cin) * len(taps),
                     f'wgrad b{b} {h}x{w} cout{cout} cin{cin} taps{len(taps)} grp{group_width}'))
  return out


def linear(x, w, bias=None, act=ACT_NONE, res=None, out_f32=False, out=None, row_map=None, res2=None,
           res2_strides=None, stats=None, drop=None):
  """x: (rows, K) bf16, w: (N, K) bf16 -> (rows, N) bf16|f32.  res: same addressing as out, added before act.
  row_map = (rows_per_group, group_stride_rows): output/res row r -> (r // rpg) * gsr + r % rpg (``out`` required)."""
  rows, k = x.shape
  n = w.shape[0]
  if row_map is None:
    a4 = x.view(1, 1, rows, k)
    st = (0, 0, n, 1)
    if out is None:
      out = torch.empty((rows, n), dtype=F32 if (out_f32 or x.dtype == F32) else BF16, device=x.device)
  else:
    rpg, gsr = row_map
    a4 = x.view(rows // rpg, 1, rpg, k)
    st = (gsr * n, 0, n, 1)
    assert out is not None
  conv_gemm(a4, w.view(n, 1, k), shift=bias, act=act, res1=res, res1_strides=st if res is not None else None, out=out,
            out_strides=st, res2=res2, res2_strides=res2_strides, stats=stats, drop=drop)
  return out


def pillar_scatter(points, use_ground_plane=False, min_x=-32.0, max_x=32.0, min_y=-32.0, max_y=32.0,
                   pixels_per_meter=4.0, hist_max=5, split_z=0.2, max_z=100.0, xform=None):
  """points (B, N, 3) f32 cuda -> (B, 1|2, 256, 256) f32 (data.py:873-906)."""
  _dev(points, F32)
  b, n, _ = points.shape
  nx = int((max_x - min_x) * pixels_per_meter)
  ny = int((max_y - min_y) * pixels_per_meter)
  counts = torch.empty((b, 2, ny, nx), dtype=torch.int32, device=points.device)
  out = torch.empty((b, 2 if use_ground_plane else 1, ny, nx), dtype=F32, device=points.device)
  if xform is not None:  # (B, n_xforms, 4) float64 {tx, ty, tz, yaw}: CARLA_Data.align fused in (data.py:840-871)
    xform = _dev(xform, torch.float64)
    check(_lib.load().tfpp_pillar_scatter_aligned(points.data_ptr(), xform.data_ptr(), xform.shape[1], b, n, counts.data_ptr(),
                                                  out.data_ptr(), int(use_ground_plane), min_x, max_x, min_y, max_y,
                                                  pixels_per_meter, hist_max, float(split_z), max_z, _stream()),
          'tfpp_pillar_scatter_aligned')
    return out
  check(_lib.load().tfpp_pillar_scatter(points.data_ptr(), b, n, counts.data_ptr(), out.data_ptr(),
                                        int(use_ground_plane), min_x, max_x, min_y, max_y, pixels_per_meter, hist_max,
                                        split_z, max_z, _stream()), 'tfpp_pillar_scatter')
  return out


def stem_conv(x, w, in_scale=None, in_shift=None, scale=None, shift=None, act=ACT_NONE, stats=None):
  _dev(x, F32)
  _dev(w, F32)
  b, cin, h, wd = x.shape
  out = torch.empty((b, h // 2, wd // 2, 32), dtype=ACT_DTYPE[0], device=x.device)
  fn = _lib.load().tfpp_stem_conv_f32 if ACT_DTYPE[0] == F32 else _lib.load().tfpp_stem_conv
  check(fn(x.data_ptr(), w.data_ptr(), _p(in_scale), _p(in_shift), _p(scale), _p(shift), act,
                                   out.data_ptr(), _p(stats[0]) if stats else None, _p(stats[1]) if stats else None, b,
     cin, h, wd, _stream()), 'tfpp_stem_conv')
  return out


def bn_finalize(stat_sum, stat_sq, gamma, beta, running_mean, running_var, count, eps=1e-5, momentum=0.1,
                save=False):
  c = stat_sum.numel()
  scale = torch.empty(c, dtype=F32, device=stat_sum.device)
  shift = torch.empty_like(scale)
  mean = torch.empty_like(scale) if save else None
  invstd = torch.empty_like(scale) if save else None
  check(_lib.load().tfpp_bn_finalize(stat_sum.data_ptr(), stat_sq.data_ptr(), _p(gamma), _p(beta), _p(running_mean),
                                     _p(running_var), scale.data_ptr(), shift.data_ptr(), _p(mean), _p(invstd), c,
                                     float(count), eps, momentum, _stream()), 'tfpp_bn_finalize')
  return scale, shift, mean, invstd


def scale_shift_act(x, scale=None, shift=None, act=ACT_NONE, res=None, pool_sum=None, out=None, res_scale=None,
                    res_shift=None):
  _dev(x)
  b, h, w, c = x.shape
  y = torch.empty_like(x) if out is None else out
  fn = _lib.load().tfpp_scale_shift_act_f32 if x.dtype == F32 else _lib.load().tfpp_scale_shift_act
  check(fn(x.data_ptr(), _p(res), _p(scale), _p(shift), _p(res_scale), _p(res_shift), act,
           y.data_ptr(), _p(pool_sum), b, h * w, c, _stream()), 'tfpp_scale_shift_act')
  return y


def se_gate(pool_sum, hw, w1, b1, w2, b2, want_hidden=False):
  b, c = pool_sum.shape
  rd = w1.shape[0]
  gate = torch.empty((b, c), dtype=F32, device=pool_sum.device)
  hidden = torch.empty((b, rd), dtype=F32, device=pool_sum.device)
  check(_lib.load().tfpp_se_gate(pool_sum.data_ptr(), hw, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                 gate.data_ptr(), _p(hidden), b, c, rd, _stream()), 'tfpp_se_gate')
  return (gate, hidden) if want_hidden else gate


def channel_scale(x, gate, out=None):
  _dev(x)
  b, h, w, c = x.shape
  y = torch.empty_like(x) if out is None else out
  fn = _lib.load().tfpp_channel_scale_f32 if x.dtype == F32 else _lib.load().tfpp_channel_scale
  check(fn(x.data_ptr(), gate.data_ptr(), y.data_ptr(), b, h * w, c, _stream()), 'tfpp_channel_scale')
  return y


def parity_split(x):
  _dev(x)
  b, h, w, c = x.shape
  y = torch.empty((4 * b, h // 2, w // 2, c), dtype=x.dtype, device=x.device)
  fn = _lib.load().tfpp_parity_split_f32 if x.dtype == F32 else _lib.load().tfpp_parity_split
  check(fn(x.data_ptr(), y.data_ptr(), b, h, w, c, _stream()), 'tfpp_parity_split')
  return y


def avgpool_tokens(x, out, ph, pw, row0, pos_emb=None):
  """x (B,H,W,C) bf16 -> rows [row0, row0+ph*pw) of out (B, rows, C) f32|bf16 (+ pos_emb (rows, C) f32)."""
  _dev(x)
  b, h, w, c = x.shape
  if x.dtype == F32:
    if out.dtype != F32:
      raise RuntimeError('fp32 mode: the token matrix must be float32')
    check(_lib.load().tfpp_avgpool_tokens_f32(x.data_ptr(), _p(pos_emb), out.data_ptr(), b, h, w, c, ph, pw, out.shape[1],
                                              row0, _stream()), 'tfpp_avgpool_tokens_f32')
    return out
  check(_lib.load().tfpp_avgpool_tokens(x.data_ptr(), _p(pos_emb), out.data_ptr(), int(out.dtype == F32), b, h, w, c,
                                        ph, pw, out.shape[1], row0, _stream()), 'tfpp_avgpool_tokens')
  return out


def bilinear(src, batch, sh, sw, dh, dw, channels, src_batch_stride=None, src_row_stride=None, add=None, src_offset=0):
  """Resize a (B,sh,sw,C) slab (f32/bf16; element strides) to NHWC bf16 (B,dh,dw,C) (+ add)."""
  if src_batch_stride is None:
    src_batch_stride = sh * sw * channels
  if src_row_stride is None:
    src_row_stride = channels
  ptr = src.data_ptr() + src_offset * src.element_size()
  if ACT_DTYPE[0] == F32:
    if src.dtype != F32 or (add is not None and add.dtype != F32):
      raise RuntimeError('fp32 mode: bilinear expects float32 tensors')
    out = torch.empty((batch, dh, dw, channels), dtype=F32, device=src.device)
    check(_lib.load().tfpp_bilinear_f32(ptr, src_batch_stride, src_row_stride, _p(add), out.data_ptr(), batch, sh, sw, dh,
                                        dw, channels, _stream()), 'tfpp_bilinear_f32')
    return out
  out = torch.empty((batch, dh, dw, channels), dtype=BF16, device=src.device)
  check(_lib.load().tfpp_bilinear(ptr, int(src.dtype == F32), src_batch_stride, src_row_stride, _p(add), out.data_ptr(),
                                  batch, sh, sw, dh, dw, channels, _stream()), 'tfpp_bilinear')
  return out


def bilinear_nchw_mask(src, channels, dh, dw, mask=None):
  _dev(src)
  b, sh, sw, cs = src.shape
  out = torch.empty((b, channels, dh, dw), dtype=F32, device=src.device)
  if src.dtype == F32:
    check(_lib.load().tfpp_bilinear_nchw_mask_f32(src.data_ptr(), _p(mask), out.data_ptr(), b, sh, sw, cs, channels, dh, dw,
                                                  _stream()), 'tfpp_bilinear_nchw_mask_f32')
    return out
  check(_lib.load().tfpp_bilinear_nchw_mask(src.data_ptr(), _p(mask), out.data_ptr(), b, sh, sw, cs, channels, dh, dw,
                                            _stream()), 'tfpp_bilinear_nchw_mask')
  return out


def nchw_to_nhwc(x):
  _dev(x, F32)
  b, c, h, w = x.shape
  y = torch.empty((b, h, w, c), dtype=BF16, device=x.device)
  check(_lib.load().tfpp_nchw_f32_to_nhwc_bf16(x.data_ptr(), y.data_ptr(), b, c, h * w, _stream()), 'nchw_to_nhwc')
  return y


def nhwc_to_nchw(x):
  _dev(x, BF16)
  b, h, w, c = x.shape
  y = torch.empty((b, c, h, w), dtype=F32, device=x.device)
  check(_lib.load().tfpp_nhwc_bf16_to_nchw_f32(x.data_ptr(), y.data_ptr(), b, c, h * w, _stream()), 'nhwc_to_nchw')
  return y


def layernorm(x, gamma, beta, want_bf16=True, want_f32=False, eps=1e-5, save=False):
  rows, c = x.shape
  if ACT_DTYPE[0] == F32:  # parity mode: the GEMM operand copy is the fp32 result itself
    want_f32 = want_f32 or want_bf16
    alias_b, want_bf16 = want_bf16, False
  else:
    alias_b = False
  yb = torch.empty((rows, c), dtype=BF16, device=x.device) if want_bf16 else None
  yf = torch.empty((rows, c), dtype=F32, device=x.device) if want_f32 else None
  mean = torch.empty(rows, dtype=F32, device=x.device) if save else None
  rstd = torch.empty(rows, dtype=F32, device=x.device) if save else None
  check(_lib.load().tfpp_layernorm(x.data_ptr(), int(x.dtype == F32), gamma.data_ptr(), beta.data_ptr(), _p(yb), _p(yf),
                                   _p(mean), _p(rstd), rows, c, eps, _stream()), 'tfpp_layernorm')
  if alias_b:
    yb = yf
  return yb, yf, mean, rstd


def _drop_args(drop):
  return (None, 0.0, 0) if drop is None else (drop[0].data_ptr(), float(drop[1]), int(drop[2]))


def dropout_(x, drop):
  """In-place dropout over a contiguous f32 / bf16 tensor (numel % 8 == 0); drop = (rng {seed, step} int64 tensor, p,
  site).  Applying it to a gradient with the same triple is the adjoint."""
  if drop is None:
    return x
  check(_lib.load().tfpp_dropout(x.data_ptr(), int(x.dtype == F32), x.numel(), *_drop_args(drop), _stream()),
        'tfpp_dropout')
  return x


def fusion_attn(qkv, batch, tokens, channels, heads, drop=None):
  _dev(qkv)
  if qkv.dtype == F32:
    out = torch.empty((batch * tokens, channels), dtype=F32, device=qkv.device)
    c3, hd = 3 * channels, channels // heads
    base = qkv.data_ptr()
    check(_lib.load().tfpp_mha_f32(base, tokens * c3, c3, base + 4 * channels, tokens * c3, c3, base + 8 * channels,
                                   tokens * c3, c3, out.data_ptr(), tokens * channels, channels, batch, heads, tokens,
                                   tokens, hd, *_drop_args(drop), _stream()), 'tfpp_mha_f32')
    return out
  out = torch.empty((batch * tokens, channels), dtype=BF16, device=qkv.device)
  check(_lib.load().tfpp_fusion_attn_dropout(qkv.data_ptr(), out.data_ptr(), batch, tokens, channels, heads,
                                             *_drop_args(drop), _stream()), 'tfpp_fusion_attn')
  return out


def small_mha(q, k, v, batch, heads, tq, tk, head_dim, q_strides, k_strides, v_strides, q_off=0, k_off=0, v_off=0,
              drop=None):
  """bf16 views given as (tensor, element offset, (batch stride, row stride)); returns (B*tq, heads*head_dim) bf16."""
  d = heads * head_dim
  if q.dtype == F32:
    out = torch.empty((batch * tq, d), dtype=F32, device=q.device)
    check(_lib.load().tfpp_mha_f32(q.data_ptr() + 4 * q_off, q_strides[0], q_strides[1], k.data_ptr() + 4 * k_off,
                                   k_strides[0], k_strides[1], v.data_ptr() + 4 * v_off, v_strides[0], v_strides[1],
                                   out.data_ptr(), tq * d, d, batch, heads, tq, tk, head_dim, *_drop_args(drop), _stream()),
          'tfpp_mha_f32')
    return out
  out = torch.empty((batch * tq, d), dtype=BF16, device=q.device)
  check(_lib.load().tfpp_small_mha_dropout(q.data_ptr() + 2 * q_off, q_strides[0], q_strides[1],
                                           k.data_ptr() + 2 * k_off, k_strides[0], k_strides[1],
                                           v.data_ptr() + 2 * v_off, v_strides[0], v_strides[1], out.data_ptr(), tq * d,
                                           d, batch, heads, tq, tk, head_dim, *_drop_args(drop), _stream()),
        'tfpp_small_mha')
  return out


def extra_sensor_token(ego_vel, command, vel_mean, vel_var, use_batch_stats, running_mean, running_var, w0, b0, w1, b1,
                       pos, mem_bf16, mem_f32, rows_per_batch, row):
  b = ego_vel.shape[0]
  check(_lib.load().tfpp_extra_sensor_token(ego_vel.data_ptr(), command.data_ptr(), float(vel_mean), float(vel_var),
                                            int(use_batch_stats), _p(running_mean), _p(running_var), w0.data_ptr(),
                                            b0.data_ptr(), w1.data_ptr(), b1.data_ptr(), pos.data_ptr(), _p(mem_bf16),
                                            _p(mem_f32), b, command.shape[1], w0.shape[0], w1.shape[0], rows_per_batch,
                                            row, _stream()), 'tfpp_extra_sensor_token')


def planner_head(joined, target_point, w_enc, b_enc, w_ih, w_hh, b_ih, b_hh, w_dec, b_dec, w_ts0, b_ts0, w_ts1, b_ts1,
                 want_h=False):
  _dev(joined, F32)
  b, nq, d = joined.shape
  n_speed = w_ts1.shape[0] if w_ts1 is not None else 0   # no target-speed token: every query feeds the GRU (wp_decoder)
  n_wp = nq - 1 if n_speed else nq
  hs = w_hh.shape[1]
  cp = torch.empty((b, n_wp, 2), dtype=F32, device=joined.device)
  ts = torch.empty((b, n_speed), dtype=F32, device=joined.device) if n_speed else None
  h_all = torch.empty((b, n_wp, hs), dtype=F32, device=joined.device) if want_h else None
  check(_lib.load().tfpp_planner_head(joined.data_ptr(), target_point.data_ptr(), w_enc.data_ptr(), b_enc.data_ptr(),
                                      w_ih.data_ptr(), w_hh.data_ptr(), b_ih.data_ptr(), b_hh.data_ptr(),
                                      w_dec.data_ptr(), b_dec.data_ptr(), _p(w_ts0), _p(b_ts0), _p(w_ts1), _p(b_ts1),
                                      cp.data_ptr(), _p(ts), _p(h_all), b, n_wp, d, hs, n_speed, _stream()),
        'tfpp_planner_head')
  return (cp, ts, h_all) if want_h else (cp, ts)


def gru_cell_head(joined, target_point, w_ih, w_hh, b_ih, b_hh, w_out, b_out, w_ts0, b_ts0, w_ts1, b_ts1, steps, hidden,
                  learn_origin=True, want_h=False):
  """GRUWaypointsPredictorTransFuser (model.py:870-913) [+ target-speed MLP]: joined (B, hidden [+2]) f32 ->
  (waypoints (B, steps, 2), speed logits (B, n_speed) | None, h_all (B, steps+1, hidden) | None)."""
  _dev(joined, F32)
  b = joined.shape[0]
  n_speed = w_ts1.shape[0] if w_ts1 is not None else 0
  wp = torch.empty((b, steps, 2), dtype=F32, device=joined.device)
  ts = torch.empty((b, n_speed), dtype=F32, device=joined.device) if n_speed else None
  h_all = torch.empty((b, steps + 1, hidden), dtype=F32, device=joined.device) if want_h else None
  check(_lib.load().tfpp_gru_cell_head(joined.data_ptr(), joined.shape[1], _p(target_point), w_ih.data_ptr(),
                                       w_hh.data_ptr(), b_ih.data_ptr(), b_hh.data_ptr(), w_out.data_ptr(), b_out.data_ptr(),
                                       _p(w_ts0), _p(b_ts0), _p(w_ts1), _p(b_ts1), wp.data_ptr(), _p(ts), _p(h_all), b, steps,
                                       hidden, w_ih.shape[1], int(learn_origin), n_speed, _stream()), 'tfpp_gru_cell_head')
  return wp, ts, h_all


def gru_cell_head_bwd(joined, target_point, w_ih, w_hh, b_ih, b_hh, w_out, b_out, w_ts0, b_ts0, w_ts1, wp, h_all, dwp, dts,
                      djoined, grads, steps, hidden, learn_origin=True):
  """BPTT of gru_cell_head; grads = (dw_ih, dw_hh, db_ih, db_hh, dw_out, db_out, dw_ts0, db_ts0, dw_ts1, db_ts1) f32
  tensors (the last four None without a target-speed head), all accumulated into, like djoined."""
  b = joined.shape[0]
  n_speed = w_ts1.shape[0] if w_ts1 is not None else 0
  check(_lib.load().tfpp_gru_cell_head_bwd(joined.data_ptr(), joined.shape[1], _p(target_point), w_ih.data_ptr(),
                                           w_hh.data_ptr(), b_ih.data_ptr(), b_hh.data_ptr(), w_out.data_ptr(),
                                           b_out.data_ptr(), _p(w_ts0), _p(b_ts0), _p(w_ts1), wp.data_ptr(), h_all.data_ptr(),
                                           dwp.data_ptr(), _p(dts), djoined.data_ptr(), *[_p(g) for g in grads], b, steps,
                                           hidden, w_ih.shape[1], int(learn_origin), n_speed, _stream()),
        'tfpp_gru_cell_head_bwd')


def decode_heatmap(heat, wh, offset, yaw_cls, yaw_res, k=100, img_h=256, img_w=256):
  """NCHW f32 maps (possibly channel-slice views with a batch stride) -> (B, k, 9) f32 (center_net.py:172-237)."""
  b, n_cls, h, w = heat.shape
  out = torch.empty((b, k, 9), dtype=F32, device=heat.device)
  for t in (heat, wh, offset, yaw_cls, yaw_res):
    if t.dtype != F32 or t.stride()[1:] != (h * w, w, 1):
      raise RuntimeError('decode_heatmap expects NCHW f32 maps with contiguous channel planes')
  check(_lib.load().tfpp_decode_heatmap(heat.data_ptr(), heat.stride(0), wh.data_ptr(), wh.stride(0), offset.data_ptr(),
                                        offset.stride(0), yaw_cls.data_ptr(), yaw_cls.stride(0), yaw_res.data_ptr(),
                                        yaw_res.stride(0), out.data_ptr(), b, n_cls, h, w, yaw_cls.shape[1], k,
                                        float(img_w / w), float(img_h / h), _stream()), 'tfpp_decode_heatmap')
  return out


def centernet_targets(boxes, counts=None, feat_h=64, feat_w=64, img_h=256, img_w=256, num_classes=4, num_dir_bins=12):
  """Ground-truth boxes (B, N <= 128, 8) f32 cuda (+ counts (B,) int32) -> the CenterNet label maps of
  CARLA_Data.get_targets (data.py:698-791) as a dict keyed like training.compute_losses' labels (+ velocity, brake)."""
  _dev(boxes, F32)
  b, n, _ = boxes.shape
  dev = boxes.device
  lab = {'center_heatmap': torch.empty((b, num_classes, feat_h, feat_w), dtype=F32, device=dev),
         'wh': torch.empty((b, 2, feat_h, feat_w), dtype=F32, device=dev),
         'offset': torch.empty((b, 2, feat_h, feat_w), dtype=F32, device=dev),
         'yaw_class': torch.empty((b, feat_h, feat_w), dtype=torch.int64, device=dev),
         'yaw_res': torch.empty((b, 1, feat_h, feat_w), dtype=F32, device=dev),
         'velocity': torch.empty((b, 1, feat_h, feat_w), dtype=F32, device=dev),
         'brake': torch.empty((b, feat_h, feat_w), dtype=torch.int64, device=dev),
         'pixel_weight': torch.empty((b, 2, feat_h, feat_w), dtype=F32, device=dev),
         'avg_factor': torch.empty(b, dtype=F32, device=dev)}
  if counts is not None:
    counts = _dev(counts, torch.int32)
  check(_lib.load().tfpp_centernet_targets(boxes.data_ptr(), _p(counts), b, n, feat_h, feat_w, img_h, img_w, num_classes,
                                           num_dir_bins, lab['center_heatmap'].data_ptr(), lab['wh'].data_ptr(),
                                           lab['offset'].data_ptr(), lab['yaw_class'].data_ptr(), lab['yaw_res'].data_ptr(),
                                           lab['velocity'].data_ptr(), lab['brake'].data_ptr(),
                                           lab['pixel_weight'].data_ptr(), lab['avg_factor'].data_ptr(), _stream()),
        'tfpp_centernet_targets')
  return lab


def nms_rotated(boxes, conf_threshold, iou_threshold, to_vehicle=False, pixels_per_meter=4.0, min_x=-32.0, min_y=-32.0,
                want_index=False):
  """boxes (B, M <= 512, S) f32 cuda, score in the last column -> (kept boxes (B, M, S) highest score first and zero
  padded, counts (B,) int32[, source rows (B, M) int32]): model.py:447-459 + transfuser_utils.py:409-452 per frame."""
  _dev(boxes, F32)
  b, m, s = boxes.shape
  out = torch.empty_like(boxes)
  count = torch.empty(b, dtype=torch.int32, device=boxes.device)
  index = torch.empty((b, m), dtype=torch.int32, device=boxes.device) if want_index else None
  check(_lib.load().tfpp_nms_rotated(boxes.data_ptr(), b, m, s, float(conf_threshold), float(iou_threshold),
                                     int(to_vehicle), float(pixels_per_meter), float(min_x), float(min_y), out.data_ptr(),
                                     count.data_ptr(), _p(index), _stream()), 'tfpp_nms_rotated')
  return (out, count, index) if want_index else (out, count)


# ---------------------------------------------------------------------------------------------- weight packing
# (load-time plumbing: layout changes of parameters, no activations involved)
def gconv3x3_supported(channels, group_width):
  return group_width == 24 and channels % 72 == 0


def gconv3x3(x, w, stride=1, scale=None, shift=None, act=ACT_NONE, stats=None):
  """Grouped 3x3 conv (group width 24, pad 1) on the haloed-tile kernel.  x (B,H,W,C) bf16, w (C/24, 9, 24, 24) bf16
  (pack_gconv_halo); returns (B,H/stride,W/stride,C) bf16.  stats = (sum, sumsq) f32 (C) accumulate the raw output."""
  _dev(x)
  _dev(w, x.dtype)
  b, h, wd, c = x.shape
  out = torch.empty((b, h // stride, wd // stride, c), dtype=x.dtype, device=x.device)
  if x.dtype == F32:
    check(_lib.load().tfpp_gconv3x3_f32(x.data_ptr(), w.data_ptr(), out.data_ptr(), _p(scale), _p(shift), act,
                                        _p(stats[0]) if stats is not None else None,
                                        _p(stats[1]) if stats is not None else None, b, h, wd, c, stride, _stream()),
          'tfpp_gconv3x3_f32')
    return out
  check(_lib.load().tfpp_gconv3x3(x.data_ptr(), w.data_ptr(), out.data_ptr(), _p(scale), _p(shift), act,
                                  _p(stats[0]) if stats is not None else None,
                                  _p(stats[1]) if stats is not None else None, b, h, wd, c, stride, _stream()),
        'tfpp_gconv3x3')
  return out


def gconv3x3_dgrad_s2(dy, w_t):
  """Input gradient of the stride-2 group conv: dy (B,Ho,Wo,C) bf16, w_t = pack_gconv_halo(w, transpose=True)."""
  _dev(dy)
  _dev(w_t, dy.dtype)
  b, ho, wo, c = dy.shape
  dx = torch.empty((b, 2 * ho, 2 * wo, c), dtype=dy.dtype, device=dy.device)
  if dy.dtype == F32:
    check(_lib.load().tfpp_gconv3x3_dgrad_s2_f32(dy.data_ptr(), w_t.data_ptr(), dx.data_ptr(), b, ho, wo, c, _stream()),
          'tfpp_gconv3x3_dgrad_s2_f32')
    return dx
  check(_lib.load().tfpp_gconv3x3_dgrad_s2(dy.data_ptr(), w_t.data_ptr(), dx.data_ptr(), b, ho, wo, c, _stream()),
        'tfpp_gconv3x3_dgrad_s2')
  return dx


def gconv3x3_wgrad(dy, x, dw, stride=1):
  """dw (C,24,3,3) f32 (torch layout, contiguous) += weight gradient of gconv3x3(x, w, stride) given dy."""
  _dev(dy)
  _dev(x, dy.dtype)
  b, h, wd, c = x.shape
  assert dw.dtype == F32 and dw.is_contiguous() and dw.numel() == c * 24 * 9
  lib = _lib.load()
  if dy.dtype == F32:
    check(lib.tfpp_gconv3x3_wgrad_f32(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), b, h, wd, c, stride, _stream()),
          'tfpp_gconv3x3_wgrad_f32')
    return dw
  ws = torch.empty(lib.tfpp_gconv3x3_wgrad_workspace(b, h, wd, c, stride), dtype=F32, device=x.device)
  check(lib.tfpp_gconv3x3_wgrad(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), ws.data_ptr(), b, h, wd, c, stride, _stream()),
        'tfpp_gconv3x3_wgrad')
  return dw


def pack_gconv_halo(w, transpose=False, dt=BF16):
  """(C, 24, 3, 3) grouped conv weight -> (C/24, 9, 24, 24) = [group][ky*3+kx][out][in]; transpose=True gives the
  input-gradient operand: in/out swapped inside each group and the taps spatially flipped."""
  c, gw, kh, kw = w.shape
  g = c // gw
  wg = w.detach().view(g, gw, gw, kh, kw)  # [g][co][ci][ky][kx]
  if transpose:
    wg = wg.flip(3, 4).permute(0, 3, 4, 2, 1)  # [g][ky'][kx'][ci][co]
  else:
    wg = wg.permute(0, 3, 4, 1, 2)  # [g][ky][kx][co][ci]
  return wg.reshape(g, kh * kw, gw, gw).to(dt).contiguous()


def gather_pack(flat, idx, out):
  """out[i] = flat[idx[i]] (0 where idx < 0), cast to out's dtype: the PackPlan refresh kernel."""
  check(_lib.load().tfpp_gather_pack(flat.data_ptr(), idx.data_ptr(), out.data_ptr(), out.numel(), int(out.dtype == F32),
                                     _stream()), 'tfpp_gather_pack')


def pack_conv_weight(w, dt=BF16):
  """(Cout, Cin, kh, kw) f32 -> (Cout, kh*kw, Cin) bf16."""
  co, ci, kh, kw = w.shape
  return w.detach().permute(0, 2, 3, 1).reshape(co, kh * kw, ci).to(dt).contiguous()


def pack_grouped_conv_weight(w, group_width=24, groups_per_tile=2, dt=BF16):
  """(Cout, gw, 3, 3) f32 grouped conv -> (Cout, 9, 64) bf16, dense within each 48-channel n-tile: column j of row n
  multiplies input channel (n // 48) * 48 + j; zero outside the row's own group."""
  co, gw, kh, kw = w.shape
  assert gw == group_width and kh == 3 and kw == 3
  tile = group_width * groups_per_tile
  out = torch.zeros((co, 9, 64), dtype=w.dtype, device=w.device)
  wt = w.detach().permute(0, 2, 3, 1).reshape(co, 9, gw)
  n = torch.arange(co, device=w.device)
  base = ((n % tile) // gw) * gw  # column offset of the row's group inside its tile
  idx = base[:, None] + torch.arange(gw, device=w.device)[None, :]
  out.scatter_(2, idx[:, None, :].expand(co, 9, gw), wt)
  return out.to(dt).contiguous()


# ---------------------------------------------------------------------------------------------- backward wrappers
TAPS_3X3_DGRAD = tuple((1 - kx, 1 - ky, 0, ky * 3 + kx) for ky in range(3) for kx in range(3))


def taps_dgrad_stride2(py, px):
  """Taps of the stride-2 3x3 dgrad for the input parity plane (py, px): input row 2u+py receives W[ky] * dY[u+dy]
  for ky == py+1 (mod 2): py=0 -> (ky=1, dy=0); py=1 -> (ky=0, dy=+1), (ky=2, dy=0)."""
  ys = ((1, 0),) if py == 0 else ((0, 1), (2, 0))
  xs = ((1, 0),) if px == 0 else ((0, 1), (2, 0))
  return tuple((dx, dy, 0, ky * 3 + kx) for ky, dy in ys for kx, dx in xs)


def pack_conv_weight_t(w, dt=BF16):
  """(Cout, Cin, kh, kw) f32 -> (Cin, kh*kw, Cout) bf16 (dgrad operand; taps not flipped, see TAPS_3X3_DGRAD)."""
  co, ci, kh, kw = w.shape
  return w.detach().permute(1, 2, 3, 0).reshape(ci, kh * kw, co).to(dt).contiguous()


def pack_grouped_conv_weight_t(w, group_width=24, dt=BF16):
  """Grouped (Cout, gw, 3, 3) -> dgrad pack (C, 9, 64): in/out swapped inside each group."""
  co, gw, kh, kw = w.shape
  g = co // gw
  wt = w.detach().view(g, gw, gw, kh, kw).permute(0, 2, 1, 3, 4).reshape(co, gw, kh, kw)
  return pack_grouped_conv_weight(wt, group_width, dt=dt)


def bn_bwd(dy, y, raw, mean, invstd, gamma, act, dgamma, dbeta, gate=None, pool_grad=None, want_dz=False,
           fwd_affine=None):
  """fwd_affine = (scale, shift) of the forward BatchNorm fold: with ReLU and y=None the mask is recomputed from raw."""
  b, h, w, c = raw.shape
  draw = torch.empty_like(raw)
  dz = torch.empty_like(raw) if want_dz else None
  fs, fh = fwd_affine if fwd_affine is not None else (None, None)
  fn = _lib.load().tfpp_bn_bwd_f32 if raw.dtype == F32 else _lib.load().tfpp_bn_bwd
  if dy.dtype != raw.dtype:
    raise RuntimeError(f'bn_bwd: dy is {dy.dtype}, raw is {raw.dtype}')
  check(fn(dy.data_ptr(), _p(y), raw.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
           gamma.data_ptr(), _p(gate), _p(pool_grad), _p(fs), _p(fh), act, dbeta.data_ptr(), dgamma.data_ptr(),
           draw.data_ptr(), _p(dz), b, h * w, c, _stream()), 'tfpp_bn_bwd')
  return draw, dz


def se_bwd(dout, a2, gate, hidden, pool_sum, hw, w1, w2, dw1, db1, dw2, db2, zeros=None):
  b, c = gate.shape
  rd = w1.shape[0]
  dgate = zeros((b, c), gate.device) if zeros else torch.zeros((b, c), dtype=F32, device=gate.device)
  pool_grad = torch.empty((b, c), dtype=F32, device=gate.device)
  ws = torch.empty((b, c + rd), dtype=F32, device=gate.device)
  dout_p, a2_p = dout.data_ptr(), a2.data_ptr()
  if dout.dtype == F32:  # fp32 parity mode: the big reduction in fp32, the (B,C)-sized rest is fp32 anyway
    check(_lib.load().tfpp_se_bwd_reduce_f32(dout_p, a2_p, dgate.data_ptr(), b, hw, c, _stream()), 'tfpp_se_bwd_reduce_f32')
    dout_p = a2_p = None
  check(_lib.load().tfpp_se_bwd(dout_p, a2_p, gate.data_ptr(), hidden.data_ptr(), pool_sum.data_ptr(),
                                hw, w1.data_ptr(), w2.data_ptr(), dgate.data_ptr(), ws.data_ptr(), dw1.data_ptr(),
                                db1.data_ptr(),
                                dw2.data_ptr(), db2.data_ptr(), pool_grad.data_ptr(), b, c, rd, _stream()),
        'tfpp_se_bwd')
  return pool_grad


def act_bwd(dy, y, act, batch, hw, channels, layout=0, act_n_limit=0, dy_scale=1.0, dbias=None, channels_padded=None,
            want_dz=True, drop=None):
  """drop = (rng, p, site) of a dropout applied to the forward output: its mask (x 1/(1-p)) multiplies dy first."""
  cp = channels if channels_padded is None else channels_padded
  if ACT_DTYPE[0] == F32:
    if dy.dtype != F32 or (y is not None and y.dtype != F32):
      raise RuntimeError('fp32 mode: act_bwd expects float32 gradients / activations')
    dz = torch.empty((batch * hw, cp), dtype=F32, device=dy.device) if want_dz else None
    check(_lib.load().tfpp_act_bwd_f32(dy.data_ptr(), _p(y), int(layout == 1), act, act_n_limit, dy_scale, _p(dz), _p(dbias),
                                       batch, hw, channels, cp, *_drop_args(drop), _stream()), 'tfpp_act_bwd_f32')
    return dz
  dz = torch.empty((batch * hw, cp), dtype=BF16, device=dy.device) if want_dz else None
  check(_lib.load().tfpp_act_bwd_dropout(dy.data_ptr(), _p(y), layout, act, act_n_limit, dy_scale, _p(dz), _p(dbias),
                                         batch, hw, channels, cp, *_drop_args(drop), _stream()), 'tfpp_act_bwd')
  return dz


def bilinear_bwd(dout, dsrc, batch, sh, sw, dh, dw, channels, src_batch_stride=None, src_row_stride=None,
                 accumulate=False, dsrc_offset=0):
  if src_batch_stride is None:
    src_batch_stride = sh * sw * channels
  if src_row_stride is None:
    src_row_stride = channels
  ptr = dsrc.data_ptr() + dsrc_offset * dsrc.element_size()
  if dout.dtype == F32:
    if dsrc.dtype != F32:
      raise RuntimeError('fp32 mode: bilinear_bwd expects a float32 destination')
    check(_lib.load().tfpp_bilinear_bwd_f32(dout.data_ptr(), ptr, src_batch_stride, src_row_stride, int(accumulate), batch, sh,
                                            sw, dh, dw, channels, _stream()), 'tfpp_bilinear_bwd_f32')
    return dsrc
  check(_lib.load().tfpp_bilinear_bwd(dout.data_ptr(), ptr, int(dsrc.dtype == F32), src_batch_stride, src_row_stride,
                                      int(accumulate), batch, sh, sw, dh, dw, channels, _stream()), 'tfpp_bilinear_bwd')
  return dsrc


def bilinear_nchw_mask_bwd(dout, mask, batch, sh, sw, src_channels, channels, dh, dw):
  if ACT_DTYPE[0] == F32:
    dsrc = torch.empty((batch, sh, sw, src_channels), dtype=F32, device=dout.device)
    check(_lib.load().tfpp_bilinear_nchw_mask_bwd_f32(dout.data_ptr(), _p(mask), dsrc.data_ptr(), batch, sh, sw, src_channels,
                                                      channels, dh, dw, _stream()), 'tfpp_bilinear_nchw_mask_bwd_f32')
    return dsrc
  dsrc = torch.empty((batch, sh, sw, src_channels), dtype=BF16, device=dout.device)
  check(_lib.load().tfpp_bilinear_nchw_mask_bwd(dout.data_ptr(), _p(mask), dsrc.data_ptr(), batch, sh, sw, src_channels,
                                                channels, dh, dw, _stream()), 'tfpp_bilinear_nchw_mask_bwd')
  return dsrc


def pool_bwd_add(dout, dtok, shape, ph, pw, rows_per_batch, row0):
  b, h, w, c = shape
  if ACT_DTYPE[0] == F32:
    if dtok.dtype != F32 or (dout is not None and dout.dtype != F32):
      raise RuntimeError('fp32 mode: pool_bwd_add expects float32 gradients')
    out = torch.empty(shape, dtype=F32, device=dtok.device)
    check(_lib.load().tfpp_pool_bwd_add_f32(_p(dout), dtok.data_ptr(), out.data_ptr(), b, h, w, c, ph, pw, rows_per_batch,
                                            row0, _stream()), 'tfpp_pool_bwd_add_f32')
    return out
  out = torch.empty(shape, dtype=BF16, device=dtok.device)
  check(_lib.load().tfpp_pool_bwd_add(_p(dout), dtok.data_ptr(), int(dtok.dtype == F32), out.data_ptr(), b, h, w, c, ph,
                                      pw, rows_per_batch, row0, _stream()), 'tfpp_pool_bwd_add')
  return out


def add_bf16(a, b, out=None):
  out = torch.empty_like(a) if out is None else out
  if a.dtype == F32:
    check(_lib.load().tfpp_add_f32(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream()), 'tfpp_add_f32')
    return out
  check(_lib.load().tfpp_add_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream()), 'tfpp_add_bf16')
  return out


def cast_rows(x, groups, group_rows, row0, rows, channels, dbias=None):
  if ACT_DTYPE[0] == F32:
    out = torch.empty((groups * rows, channels), dtype=F32, device=x.device)
    check(_lib.load().tfpp_copy_rows_f32(x.data_ptr(), out.data_ptr(), _p(dbias), groups, group_rows, row0, rows, channels,
                                         _stream()), 'tfpp_copy_rows_f32')
    return out
  out = torch.empty((groups * rows, channels), dtype=BF16, device=x.device)
  check(_lib.load().tfpp_cast_rows(x.data_ptr(), out.data_ptr(), _p(dbias), groups, group_rows, row0, rows, channels,
                                   _stream()), 'tfpp_cast_rows')
  return out


def batch_reduce(x, out, batch):
  check(_lib.load().tfpp_batch_reduce(x.data_ptr(), out.data_ptr(), batch, out.numel(), _stream()), 'tfpp_batch_reduce')


def stem_wgrad(x, draw, in_scale, in_shift, dw):
  b, cin, h, w = x.shape
  if draw.dtype == F32:
    check(_lib.load().tfpp_stem_wgrad_f32(x.data_ptr(), draw.data_ptr(), _p(in_scale), _p(in_shift), dw.data_ptr(), b, cin, h,
                                          w, _stream()), 'tfpp_stem_wgrad_f32')
    return
  check(_lib.load().tfpp_stem_wgrad(x.data_ptr(), draw.data_ptr(), _p(in_scale), _p(in_shift), dw.data_ptr(), b, cin, h,
                                    w, _stream()), 'tfpp_stem_wgrad')


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, dres=None):
  rows, c = x.shape
  dx = torch.empty((rows, c), dtype=F32, device=x.device)
  check(_lib.load().tfpp_layernorm_bwd(dy.data_ptr(), int(dy.dtype == F32), x.data_ptr(), mean.data_ptr(),
                                       rstd.data_ptr(), gamma.data_ptr(), _p(dres), dx.data_ptr(), dgamma.data_ptr(),
                                       dbeta.data_ptr(), rows, c, _stream()), 'tfpp_layernorm_bwd')
  return dx


def fusion_attn_bwd(qkv, dout, batch, tokens, channels, heads, drop=None):
  dqkv = torch.empty_like(qkv)
  if qkv.dtype == F32:
    c3, hd = 3 * channels, channels // heads
    ws = torch.empty(2 * batch * heads * tokens * tokens, dtype=F32, device=qkv.device)
    base, dbase = qkv.data_ptr(), dqkv.data_ptr()
    sb, sr = tokens * c3, c3
    check(_lib.load().tfpp_mha_bwd_f32(base, sb, sr, base + 4 * channels, sb, sr, base + 8 * channels, sb, sr, dout.data_ptr(),
                                       tokens * channels, channels, dbase, sb, sr, dbase + 4 * channels, sb, sr,
                                       dbase + 8 * channels, sb, sr, ws.data_ptr(), 0, batch, heads, tokens, tokens, hd,
                                       *_drop_args(drop), _stream()), 'tfpp_mha_bwd_f32')
    return dqkv
  ws = torch.empty((batch * tokens, 2 * channels), dtype=F32, device=qkv.device)
  check(_lib.load().tfpp_fusion_attn_bwd_dropout(qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), ws.data_ptr(), batch,
                                                 tokens, channels, heads, *_drop_args(drop), _stream()),
        'tfpp_fusion_attn_bwd')
  return dqkv


def small_mha_bwd(q, k, v, dout, dq, dk, dv, batch, heads, tq, tk, head_dim, q_st, k_st, v_st, dq_st, dk_st, dv_st,
                  offs=(0, 0, 0, 0, 0, 0), accumulate_kv=False, drop=None):
  """offs: element offsets of (q, k, v, dq, dk, dv) inside their buffers; *_st = (batch stride, row stride)."""
  d = heads * head_dim
  if q.dtype == F32:
    p4 = lambda t, o: t.data_ptr() + 4 * o
    ws = torch.empty(2 * batch * heads * tq * tk, dtype=F32, device=q.device)
    check(_lib.load().tfpp_mha_bwd_f32(p4(q, offs[0]), q_st[0], q_st[1], p4(k, offs[1]), k_st[0], k_st[1], p4(v, offs[2]),
                                       v_st[0], v_st[1], dout.data_ptr(), tq * d, d, p4(dq, offs[3]), dq_st[0], dq_st[1],
                                       p4(dk, offs[4]), dk_st[0], dk_st[1], p4(dv, offs[5]), dv_st[0], dv_st[1],
                                       ws.data_ptr(), int(accumulate_kv), batch, heads, tq, tk, head_dim, *_drop_args(drop),
                                       _stream()), 'tfpp_mha_bwd_f32')
    return
  p = lambda t, o: t.data_ptr() + 2 * o
  check(_lib.load().tfpp_small_mha_bwd_dropout(p(q, offs[0]), q_st[0], q_st[1], p(k, offs[1]), k_st[0], k_st[1],
                                               p(v, offs[2]), v_st[0], v_st[1], dout.data_ptr(), tq * d, d,
                                               p(dq, offs[3]), dq_st[0], dq_st[1], p(dk, offs[4]), dk_st[0], dk_st[1],
                                               p(dv, offs[5]), dv_st[0], dv_st[1], int(accumulate_kv), batch, heads, tq,
                                               tk, head_dim, *_drop_args(drop), _stream()), 'tfpp_small_mha_bwd')


# ---------------------------------------------------------------------------------------------- experimental: halo UMMA
def pack_halo_umma_weight(w, n_pad, transpose=False, k_pad=None, dt=BF16):
  """(Cout, Cin, 3, 3) -> (9, K/8, n_pad, 8) = [tap][k chunk][n][8 k], the non-swizzled K-major UMMA operand layout
  (zero padded to n_pad output and k_pad input channels).  transpose=True: the input-gradient operand (in/out
  swapped, taps spatially flipped)."""
  wd = w.detach()
  if transpose:
    wd = wd.flip(2, 3).permute(1, 0, 2, 3)
  n, k = wd.shape[0], wd.shape[1]
  k_pad = k if k_pad is None else k_pad
  assert k_pad % 8 == 0 and k <= k_pad and n <= n_pad
  if k_pad > k or n_pad > n:
    wd = torch.nn.functional.pad(wd, (0, 0, 0, 0, 0, k_pad - k, 0, n_pad - n))
  t = wd.permute(2, 3, 1, 0).reshape(9, k_pad // 8, 8, n_pad).permute(0, 1, 3, 2)  # [tap][kc][n][8]
  return t.to(dt).contiguous()


def halo_umma_supported(cin, n_pad):
  return cin in (16, 32, 64) and n_pad in (16, 32, 48, 64) and 9 * cin * n_pad * 2 + 2 * (cin // 8) * 10368 <= 220 * 1024


def pack_halo_gconv_weight(w, transpose=False, dt=BF16):
  """(C, 24, 3, 3) group conv weight -> (C/24, 9, 4, 32, 8) = [group][tap][k chunk][n][8 k], zero padded from 24 to 32
  input and output channels (tfpp_halo_gconv3x3).  transpose=True: the stride-1 input-gradient operand."""
  c, gw, kh, kw = w.shape
  g = c // gw
  wg = w.detach().view(g, gw, gw, kh, kw)  # [g][co][ci][ky][kx]
  if transpose:
    wg = wg.flip(3, 4).permute(0, 2, 1, 3, 4)  # [g][ci (new out)][co (new in)][ky'][kx']
  wg = torch.nn.functional.pad(wg, (0, 0, 0, 0, 0, 32 - gw, 0, 32 - gw))  # [g][n 32][k 32][3][3]
  t = wg.permute(0, 3, 4, 2, 1).reshape(g, 9, 4, 8, 32).permute(0, 1, 2, 4, 3)  # [g][tap][kc][n][8]
  return t.to(dt).contiguous()


def halo_gconv3x3(x, w_packed, scale=None, shift=None, act=ACT_NONE, stats=None):
  """EXPERIMENTAL tcgen05 group conv (stride 1) over haloed planes; same contract as gconv3x3(x, w, 1, ...)."""
  _dev(x, BF16)
  _dev(w_packed, BF16)
  b, h, wd, c = x.shape
  out = torch.empty((b, h, wd, c), dtype=BF16, device=x.device)
  check(_lib.load().tfpp_halo_gconv3x3(x.data_ptr(), w_packed.data_ptr(), out.data_ptr(), _p(scale), _p(shift), act,
                                       _p(stats[0]) if stats is not None else None,
                                       _p(stats[1]) if stats is not None else None, b, h, wd, c, _stream()),
        'tfpp_halo_gconv3x3')
  return out


def halo_conv3x3(x, w_packed, bias=None, act=ACT_NONE, act_n_limit=0, n_valid=None, out_nchw_f32=False):
  """EXPERIMENTAL tcgen05 haloed-tile 3x3 conv (csrc/halo_umma.cu); same contract as smallc_conv3x3."""
  _dev(x, BF16)
  _dev(w_packed, BF16)
  b, h, wd, cin = x.shape
  n_pad = w_packed.shape[2]
  n_valid = n_pad if n_valid is None else n_valid
  if out_nchw_f32:
    out = torch.empty((b, n_valid, h, wd), dtype=F32, device=x.device)
  else:
    out = torch.empty((b, h, wd, n_pad), dtype=BF16, device=x.device)
  check(_lib.load().tfpp_halo_conv3x3(x.data_ptr(), w_packed.data_ptr(), _p(bias), out.data_ptr(), int(out_nchw_f32),
                                      n_valid, act, act_n_limit, b, h, wd, cin, n_pad, _stream()), 'tfpp_halo_conv3x3')
  return out


# ---------------------------------------------------------------------------------------------- small-channel convs
def smallc_supported(cin, cout_pad):
  return (cin, cout_pad) in ((32, 32), (32, 16), (32, 8), (16, 32))


def smallc_conv3x3(x, w, bias=None, act=ACT_NONE, act_n_limit=0, n_valid=None, out_nchw_f32=False):
  """x (B,H,W,Cin) bf16, w (Cout_pad, 9, Cin) bf16 -> NHWC bf16 (B,H,W,Cout_pad) or NCHW f32 (B,n_valid,H,W)."""
  _dev(x, BF16)
  _dev(w, BF16)
  b, h, wd, cin = x.shape
  cout = w.shape[0]
  n_valid = cout if n_valid is None else n_valid
  if out_nchw_f32:
    out = torch.empty((b, n_valid, h, wd), dtype=F32, device=x.device)
  else:
    out = torch.empty((b, h, wd, cout), dtype=BF16, device=x.device)
  check(_lib.load().tfpp_smallc_conv3x3(x.data_ptr(), w.data_ptr(), _p(bias), out.data_ptr(), int(out_nchw_f32), n_valid,
                                        act, act_n_limit, b, h, wd, cin, cout, _stream()), 'tfpp_smallc_conv3x3')
  return out


def smallc_wgrad3x3(dy, x, out, out_strides, co_valid):
  _dev(dy, BF16)
  _dev(x, BF16)
  b, h, w, cop = dy.shape
  check(_lib.load().tfpp_smallc_wgrad3x3(dy.data_ptr(), x.data_ptr(), out.data_ptr(), out_strides[0], out_strides[1],
                                         out_strides[2], co_valid, b, h, w, cop, x.shape[3], _stream()),
        'tfpp_smallc_wgrad3x3')
  return out


# ---------------------------------------------------------------------------------------------- bev_encoder backbone
def instnorm(x, act=ACT_NONE, eps=1e-5, out=None, out_pix_stride=None, zeros=None, save=False):
  """nn.InstanceNorm2d(affine=False) + activation on an NHWC (B,H,W,C) bf16|f32 tensor.  ``out`` may be a wider NHWC
  tensor (out_pix_stride elements per pixel, the C outputs go to its first C channels).  Returns (y, mean, invstd)."""
  _dev(x)
  b, h, w, c = x.shape
  f32 = int(x.dtype == F32)
  mk = zeros if zeros is not None else (lambda shape, dev: torch.zeros(shape, dtype=F32, device=dev))
  sums = mk((2, b, c), x.device)
  check(_lib.load().tfpp_instnorm_stats(x.data_ptr(), f32, c, b, h * w, c, sums[0].data_ptr(), sums[1].data_ptr(), _stream()),
        'tfpp_instnorm_stats')
  if out is None:
    out = torch.empty_like(x)
    out_pix_stride = c
  else:
    if out.dtype != x.dtype:
      raise RuntimeError(f'instnorm: x is {x.dtype}, out is {out.dtype}')
    out_pix_stride = out_pix_stride or out.shape[-1]
  mean = torch.empty((b, c), dtype=F32, device=x.device) if save else None
  invstd = torch.empty((b, c), dtype=F32, device=x.device) if save else None
  check(_lib.load().tfpp_instnorm_apply(x.data_ptr(), f32, c, sums[0].data_ptr(), sums[1].data_ptr(), eps, act, out.data_ptr(),
                                        out_pix_stride, _p(mean), _p(invstd), b, h * w, c, _stream()), 'tfpp_instnorm_apply')
  return out, mean, invstd


def instnorm_bwd(dy, x, mean, invstd, act, dy_pix_stride=None, zeros=None):
  """Adjoint of instnorm wrt x.  dy: gradient wrt the activation output, NHWC with dy_pix_stride elements per pixel."""
  _dev(x)
  b, h, w, c = x.shape
  if dy.dtype != x.dtype:
    raise RuntimeError(f'instnorm_bwd: dy is {dy.dtype}, x is {x.dtype}')
  mk = zeros if zeros is not None else (lambda shape, dev: torch.zeros(shape, dtype=F32, device=dev))
  s = mk((2, b, c), x.device)
  dx = torch.empty_like(x)
  check(_lib.load().tfpp_instnorm_bwd(dy.data_ptr(), dy_pix_stride or c, x.data_ptr(), int(x.dtype == F32), mean.data_ptr(),
                                      invstd.data_ptr(), act, s[0].data_ptr(), s[1].data_ptr(), dx.data_ptr(), b, h * w, c,
                                      _stream()), 'tfpp_instnorm_bwd')
  return dx


def bev_lift(img, tables, depth, width):
  """img (B,IH,IW,C) NHWC -> (B, width, depth, C) NHWC; tables = (a_rows (depth, IH) f32, x0 (depth, width) int32,
  wl, wr (depth, width) f32) from nn.bev_encoder.lift_tables."""
  _dev(img)
  b, ih, iw, c = img.shape
  a, x0, wl, wr = tables
  out = torch.empty((b, width, depth, c), dtype=img.dtype, device=img.device)
  check(_lib.load().tfpp_bev_lift(img.data_ptr(), int(img.dtype == F32), a.data_ptr(), x0.data_ptr(), wl.data_ptr(),
                                  wr.data_ptr(), out.data_ptr(), b, ih, iw, c, depth, width, _stream()), 'tfpp_bev_lift')
  return out


def bev_lift_bwd(dout, tables, img_shape, dimg=None):
  """Adjoint of bev_lift wrt img; accumulates into ``dimg`` when given."""
  _dev(dout)
  b, ih, iw, c = img_shape
  _, width, depth, _ = dout.shape
  a, x0, wl, wr = tables
  ws = torch.empty((b, depth, iw, c), dtype=F32, device=dout.device)
  acc = dimg is not None
  if dimg is None:
    dimg = torch.empty(img_shape, dtype=dout.dtype, device=dout.device)
  check(_lib.load().tfpp_bev_lift_bwd(dout.data_ptr(), int(dout.dtype == F32), a.data_ptr(), x0.data_ptr(), wl.data_ptr(),
                                      wr.data_ptr(), ws.data_ptr(), dimg.data_ptr(), int(acc), b, ih, iw, c, depth, width,
                                      _stream()), 'tfpp_bev_lift_bwd')
  return dimg
