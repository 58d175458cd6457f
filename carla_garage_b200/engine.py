"""Forward (and, in engine_bwd, backward) schedule of the TransFuser++ step on the libtfpp.so kernels.

The nn.Module tree (carla_garage_b200.nn) only holds parameters in the reference's layout; this file is the
B200-native execution plan: NHWC bf16 feature maps, an fp32 (B,320,C) token residual stream per fusion scale,
tcgen05 implicit-GEMM for every convolution / projection with BatchNorm statistics, bias, activation and residuals
in the epilogue, and small dedicated kernels for everything else.  Each step cites the reference lines it replaces
(paths relative to /root/reference/team_code).
"""
import os
import weakref

import torch

from . import ops
from .ops import ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_GELU, BF16, F32

_PACK_CACHE = {}
HALO_UMMA = os.environ.get('TFPP_HALO_UMMA', '0') == '1'  # experimental tcgen05 haloed-tile convs (round 2)
HALO_UMMA_GCONV = os.environ.get('TFPP_HALO_UMMA_GCONV', '0') == '1'  # same for the RegNet group convs (never run yet)
PARAM_EPOCH = [0]  # bumped by the fused optimizer (it updates parameter storage without touching version counters)


# kinds whose pack is a pure gather (+ zero padding) of parameter elements: eligible for the one-kernel PackPlan
_GATHER_KINDS = frozenset(('conv', 'gconv', 'gconv_halo', 'gconv_halo_t', 'gconv_halo_umma', 'gconv_halo_umma_t', 'conv_halo_umma', 'conv_halo_umma_t', 'linear', 'conv_t', 'conv_rows_pad', 'conv_dgrad_smallc', 'gconv_t', 'linear_t',
                           'rows_t', 'cols', 'cols_t', 'conv_cin', 'conv_cin_t', 'conv_cin_pad', 'conv_cin_pad_t', 'cat_linear_t', 'cat_conv_t', 'blockdiag_1x1_t', 'rows', 'rows_f32', 'cat_linear',
                           'cat_rows', 'cat_rows_f32', 'cat_f32', 'cat_conv', 'blockdiag_1x1', 'repeat_rows'))


def _build_pack(kind, params, extra, dtb=BF16, dtf=F32):
  """The kernel-layout tensor(s) of one pack.  dtb / dtf are the element types of the bf16 / fp32 results; PackPlan
  calls this a second time with float64 'index tensors' to learn where every packed element comes from."""
  if kind == 'conv':  # (Cout,Cin,kh,kw) -> (Cout, kh*kw, Cin) bf16
    return ops.pack_conv_weight(params[0], dt=dtb)
  if kind == 'gconv':
    return ops.pack_grouped_conv_weight(params[0], dt=dtb)
  if kind == 'gconv_halo_umma':  # experimental tcgen05 group conv: (C/24, 9, 4, 32, 8)
    return ops.pack_halo_gconv_weight(params[0], dt=dtb)
  if kind == 'gconv_halo_umma_t':
    return ops.pack_halo_gconv_weight(params[0], transpose=True, dt=dtb)
  if kind == 'conv_halo_umma':  # experimental tcgen05 haloed-tile conv: (9, Cin/8, extra[0], 8)
    return ops.pack_halo_umma_weight(params[0], extra[0], dt=dtb)
  if kind == 'conv_halo_umma_t':  # its input-gradient operand, K padded to extra[0]
    return ops.pack_halo_umma_weight(params[0], params[0].shape[1], transpose=True, k_pad=extra[0], dt=dtb)
  if kind == 'gconv_halo':  # (C,24,3,3) -> (C/24, 9, 24, 24) for tfpp_gconv3x3
    return ops.pack_gconv_halo(params[0], dt=dtb)
  if kind == 'gconv_halo_t':  # its input-gradient operand
    return ops.pack_gconv_halo(params[0], transpose=True, dt=dtb)
  if kind == 'linear':
    return params[0].detach().reshape(params[0].shape[0], -1).to(dtb).contiguous()
  if kind == 'conv_t':  # dgrad operand (Cin, taps, Cout padded to a multiple of 8)
    w = ops.pack_conv_weight_t(params[0], dt=dtb)
    pad = (-w.shape[2]) % 8
    return torch.nn.functional.pad(w, (0, pad)).contiguous() if pad else w
  if kind == 'conv_rows_pad':  # (Cout,Cin,3,3) -> (Cout padded to extra[0], 9, Cin) bf16
    w = ops.pack_conv_weight(params[0], dt=dtb)
    return torch.nn.functional.pad(w, (0, 0, 0, 0, 0, extra[0] - w.shape[0])).contiguous()
  if kind == 'conv_dgrad_smallc':  # (Cout,Cin,3,3) -> (Cin, 9 flipped taps, Cout padded to extra[0]) bf16
    w = params[0].detach().flip(2, 3).permute(1, 2, 3, 0).reshape(params[0].shape[1], 9, params[0].shape[0])
    return torch.nn.functional.pad(w, (0, extra[0] - w.shape[2])).to(dtb).contiguous()
  if kind == 'gconv_t':
    return ops.pack_grouped_conv_weight_t(params[0], dt=dtb)
  if kind == 'linear_t':  # (N,K) -> (K, N padded to 8)
    w = params[0].detach().reshape(params[0].shape[0], -1).t().to(dtb).contiguous()
    pad = (-w.shape[1]) % 8
    return torch.nn.functional.pad(w, (0, pad)).contiguous() if pad else w
  if kind == 'cols':  # column slice of a (N,K) matrix: the part of a Linear that multiplies one piece of a concatenated input
    return params[0].detach()[:, extra[0]:extra[1]].to(dtb).contiguous()
  if kind == 'cols_t':  # its input-gradient operand: (k1-k0, N padded to 8)
    w = params[0].detach()[:, extra[0]:extra[1]].t().to(dtb).contiguous()
    pad = (-w.shape[1]) % 8
    return torch.nn.functional.pad(w, (0, pad)).contiguous() if pad else w
  if kind == 'conv_cin':  # input-channel slice [k0, k1) of a conv weight: (Cout, taps, k1 - k0) — one piece of a conv over a concatenation
    return ops.pack_conv_weight(params[0].detach()[:, extra[0]:extra[1]], dt=dtb)
  if kind == 'conv_cin_t':  # its input-gradient operand (k1 - k0, taps, Cout padded to 8)
    w = ops.pack_conv_weight_t(params[0].detach()[:, extra[0]:extra[1]], dt=dtb)
    pad = (-w.shape[2]) % 8
    return torch.nn.functional.pad(w, (0, pad)).contiguous() if pad else w
  if kind == 'conv_cin_pad':  # (Cout, Cin, kh, kw) -> (Cout, taps, Cin zero-padded to extra[0])
    w = ops.pack_conv_weight(params[0], dt=dtb)
    return torch.nn.functional.pad(w, (0, extra[0] - w.shape[2])).contiguous()
  if kind == 'conv_cin_pad_t':  # its input-gradient operand (Cin padded to extra[0], taps, Cout)
    w = ops.pack_conv_weight_t(params[0], dt=dtb)
    return torch.nn.functional.pad(w, (0, 0, 0, 0, 0, extra[0] - w.shape[0])).contiguous()
  if kind == 'rows_t':
    return params[0].detach()[extra[0]:extra[1]].t().to(dtb).contiguous()
  if kind == 'cat_linear_t':
    return torch.cat([p.detach().reshape(p.shape[0], -1) for p in params], dim=0).t().to(dtb).contiguous()
  if kind == 'cat_conv_t':  # several convs stacked along Cout -> (Cin, taps, sum Cout)
    return torch.cat([ops.pack_conv_weight_t(p, dt=dtb) for p in params], dim=2).contiguous()
  if kind == 'blockdiag_1x1_t':  # transpose of blockdiag_1x1: (sum K_i, 1, sum N_i padded to 8)
    n = sum(p.shape[0] for p in params)
    k = sum(p.shape[1] for p in params)
    npad = n + ((-n) % 8)
    out = torch.zeros((k, 1, npad), dtype=params[0].dtype, device=params[0].device)
    r = c = 0
    for p in params:
      out[c:c + p.shape[1], 0, r:r + p.shape[0]] = p.detach().reshape(p.shape[0], p.shape[1]).t()
      r += p.shape[0]
      c += p.shape[1]
    return out.to(dtb).contiguous()
  if kind == 'rows':  # row slice of a (N,K) matrix
    return params[0].detach()[extra[0]:extra[1]].to(dtb).contiguous()
  if kind == 'rows_f32':
    return params[0].detach()[extra[0]:extra[1]].to(dtf).contiguous()
  if kind == 'cat_linear':
    return torch.cat([p.detach().reshape(p.shape[0], -1) for p in params], dim=0).to(dtb).contiguous()
  if kind == 'cat_rows':  # rows [r0,r1) of several matrices stacked
    return torch.cat([p.detach()[extra[0]:extra[1]] for p in params], dim=0).to(dtb).contiguous()
  if kind == 'cat_rows_f32':
    return torch.cat([p.detach()[extra[0]:extra[1]] for p in params], dim=0).to(dtf).contiguous()
  if kind == 'cat_f32':
    return torch.cat([p.detach().reshape(-1) for p in params], dim=0).to(dtf).contiguous()
  if kind == 'cat_conv':
    return torch.cat([ops.pack_conv_weight(p, dt=dtb) for p in params], dim=0).contiguous()
  if kind == 'blockdiag_1x1':
    n = sum(p.shape[0] for p in params)
    k = sum(p.shape[1] for p in params)
    out = torch.zeros((n, 1, k), dtype=params[0].dtype, device=params[0].device)
    r = c = 0
    for p in params:
      out[r:r + p.shape[0], 0, c:c + p.shape[1]] = p.detach().reshape(p.shape[0], p.shape[1])
      r += p.shape[0]
      c += p.shape[1]
    return out.to(dtb).contiguous()
  if kind == 'bn_eval':  # (gamma, beta, running_mean, running_var) -> folded (scale, shift)
    g, b, m, v = (p.detach().float() for p in params)
    scale = g * torch.rsqrt(v + extra[0])
    return (scale.contiguous(), (b - m * scale).contiguous())
  if kind == 'f32':
    return params[0].detach().float().contiguous()
  if kind == 'host_floats':  # one device->host read, cached until the buffers change
    return tuple(float(p.detach().reshape(-1)[0]) for p in params)
  if kind == 'repeat_rows':  # (1, T, C) parameter repeated over the batch: f32 and bf16 copies
    t = params[0].detach().to(dtf).reshape(-1, params[0].shape[-1]).repeat(extra[0], 1).contiguous()
    return (t, t.to(dtb))
  raise ValueError(kind)


class PackPlan:
  """All gather-type weight packs of a training step as ONE kernel.

  The parameters of a Trainer live in one flat fp32 buffer (training.FlatState).  The first (eager) step builds every
  pack with torch ops as usual and, next to it, an int32 map 'packed element -> flat parameter index' (-1 = zero
  padding), obtained by running the same pack code on index tensors.  finalize() concatenates the maps; from then on
  packed() hands out fixed views of one bf16 (and one fp32) buffer, and refresh() — one tfpp_gather_pack launch per
  buffer, called right after the optimizer — rewrites all of them.  That replaces ~2000 small cast / permute / pad
  kernels per step and gives the CUDA graph static weight addresses."""

  ALIGN = 128  # elements; keeps every region 256-byte aligned for TMA descriptors

  def __init__(self, flat):
    self.flat = flat
    self.lo = flat.data_ptr()
    self.hi = self.lo + flat.numel() * 4
    self.views = {}     # key -> (versions, out structure of views, params, kind, extra)
    self.pending = {}   # key -> (params, [(idx int32 tensor, like tensor), ...], is_tuple)
    self.segments = []  # (idx_all, out_all, is_f32)

  def offset_of(self, p):
    ptr = p.data_ptr()
    if p.dtype != F32 or not p.is_contiguous() or ptr < self.lo or ptr + p.numel() * 4 > self.hi:
      return None
    return (ptr - self.lo) // 4

  def lookup(self, key, params):
    hit = self.views.get(key)
    if hit is None:
      return None
    ver = tuple(p._version for p in params)  # pylint: disable=protected-access
    if hit[0] != ver:  # somebody wrote the parameters with torch ops (load_state_dict, tests): re-gather everything
      self.refresh()
      self.views[key] = (ver,) + hit[1:]
    return hit[1]

  def register(self, key, kind, params, extra, out):
    if kind not in _GATHER_KINDS or key in self.pending:
      return
    offs = [self.offset_of(p) for p in params]
    if any(o is None for o in offs):
      return
    probes = [(torch.arange(p.numel(), dtype=torch.float64, device=p.device) + (o + 1)).view(p.shape)
              for p, o in zip(params, offs)]
    idx = _build_pack(kind, probes, extra, dtb=torch.float64, dtf=torch.float64)
    is_tuple = isinstance(out, tuple)
    outs = out if is_tuple else (out,)
    idxs = idx if is_tuple else (idx,)
    items = []
    for o, i in zip(outs, idxs):
      assert o.shape == i.shape and o.dtype in (BF16, F32)
      items.append(((i.reshape(-1) - 1).to(torch.int32), o))
    self.pending[key] = (tuple(params), items, is_tuple, kind, extra)

  def finalize(self):
    """Move every pack registered since the last call into plan-owned storage."""
    if not self.pending:
      return
    dev = self.flat.device
    for is_f32 in (False, True):
      dt = F32 if is_f32 else BF16
      total = 0
      places = []
      for key, (_, items, _, _, _) in self.pending.items():
        for j, (idx, like) in enumerate(items):
          if like.dtype == dt:
            places.append((key, j, total, idx, like))
            total += -(-idx.numel() // self.ALIGN) * self.ALIGN
      if not total:
        continue
      idx_all = torch.full((total,), -1, dtype=torch.int32, device=dev)
      out_all = torch.empty((total,), dtype=dt, device=dev)
      for key, j, off, idx, like in places:
        idx_all[off:off + idx.numel()] = idx
        self.pending[key][1][j] = (None, out_all[off:off + idx.numel()].view(like.shape))
      self.segments.append((idx_all, out_all, is_f32))
      ops.gather_pack(self.flat, idx_all, out_all)
    for key, (params, items, is_tuple, kind, extra) in self.pending.items():
      outs = tuple(v for _, v in items)
      ver = tuple(p._version for p in params)  # pylint: disable=protected-access
      self.views[key] = (ver, outs if is_tuple else outs[0], params, kind, extra)
      _PACK_CACHE.pop(key, None)
    self.pending = {}

  def refresh(self):
    for idx_all, out_all, _ in self.segments:
      ops.gather_pack(self.flat, idx_all, out_all)

  def refresh_all(self):
    """One re-gather after the parameters were written with torch ops (torch optimizer, load_state_dict): every view
    adopts the current version counters, so lookup() does not re-gather once per pack."""
    self.refresh()
    for key, hit in self.views.items():
      self.views[key] = (tuple(p._version for p in hit[2]),) + hit[1:]  # pylint: disable=protected-access


_PLAN = [None]  # the active PackPlan (set by training.Trainer)


def packed(params, kind, *extra):
  """Kernel-layout copy of parameter(s); rebuilt whenever a version counter / storage changes (optimizer step,
  load_state_dict, .to(device)) — or, under a PackPlan, a fixed view that PackPlan.refresh() keeps current."""
  params = params if isinstance(params, (tuple, list)) else (params,)
  dtb = ops.act_dtype()  # bf16 packs in production, fp32 packs in the parity mode (ops.set_precision('fp32'))
  key = (kind,) + tuple(id(p) for p in params) + extra + ((dtb,) if dtb != BF16 else ())
  plan = _PLAN[0]
  if plan is not None:
    hit = plan.lookup(key, params)
    if hit is not None:
      return hit
  ver = (PARAM_EPOCH[0],) + tuple(p._version for p in params) + tuple(p.data_ptr() for p in params)  # pylint: disable=protected-access
  hit = _PACK_CACHE.get(key)
  # the key holds id()s: a hit only counts while the very same parameter objects are alive (a freed model's ids and
  # addresses can be handed to a new one)
  if hit is not None and hit[0] == ver and all(r() is p for r, p in zip(hit[2], params)):
    return hit[1]
  with torch.no_grad():
    out = _build_pack(kind, params, extra, dtb=dtb)
    if plan is not None and dtb == BF16:
      plan.register(key, kind, params, extra, out)
  _PACK_CACHE[key] = (ver, out, tuple(weakref.ref(p) for p in params))
  return out


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class Engine:
  """Execution plan bound to one LidarCenterNet (or a bare TransfuserBackbone / LidarCenterNetHead)."""

  def __init__(self, model=None, backbone=None, head=None):
    self.m = model
    self.bb = backbone if backbone is not None else (model.backbone if model is not None else None)
    self.head = head if head is not None else (getattr(model, 'head', None) if model is not None else None)
    self.cfg = (model or backbone or head).config
    self._consts = {}
    self.tape = None  # list of saved activations when a backward pass will follow
    self.debug_taps = None  # dict: name -> NHWC bf16 intermediate (tests only)
    # training-mode dropout (transfuser.py:325,374,379,395; nn.TransformerDecoderLayer / MultiheadAttention 0.1):
    # counter-based masks keyed by (seed, step, site, element) — see tfpp_dropout in include/tfpp.h
    self.dropout_enabled = os.environ.get('TFPP_DROPOUT', '1') != '0'
    self.rng = None      # int64 device tensor {seed, step}; step is bumped once per training forward
    self._site = 0       # dropout sites are numbered in call order within one forward

  @classmethod
  def for_backbone(cls, backbone):
    eng = getattr(backbone, '_tfpp_engine', None)
    if eng is None:
      eng = cls(backbone=backbone)
      object.__setattr__(backbone, '_tfpp_engine', eng)
    return eng

  @classmethod
  def for_head(cls, head):
    eng = getattr(head, '_tfpp_engine', None)
    if eng is None:
      eng = cls(head=head)
      object.__setattr__(head, '_tfpp_engine', eng)
    return eng

  # ------------------------------------------------------------------------------------------------ helpers
  def _const(self, name, fn, device):
    key = (name, str(device))
    if key not in self._consts:
      self._consts[key] = fn().to(device)
    return self._consts[key]

  def _save(self, **kw):
    if self.tape is not None:
      self.tape.append(kw)

  def side_stream(self, device, which='planner'):
    key = ('side_stream', which, str(device))
    if key not in self._consts:
      self._consts[key] = torch.cuda.Stream(device=device)
    return self._consts[key]

  def seed_dropout(self, seed, device=None, step=0):
    """(Re)seed the dropout stream: masks are a pure function of (seed, step, site, element index)."""
    device = device or (self.rng.device if self.rng is not None else 'cuda')
    self.rng = torch.tensor([int(seed) & (2**63 - 1), int(step)], dtype=torch.int64, device=device)

  def begin_dropout_step(self, device):
    """Start of a training forward: next step of the random stream, site numbering restarts."""
    device = torch.device(device)
    if device.type == 'cuda' and device.index is None:
      device = torch.device('cuda', torch.cuda.current_device())   # 'cuda' and 'cuda:0' are the same place
    if self.rng is None or self.rng.device != device:
      self.seed_dropout(torch.initial_seed(), device)
    self.rng[1:2] += 1
    self._site = 0

  def drop(self, p, training=True):
    """(rng, p, site) of the next dropout site, or None when dropout is inactive."""
    if not (training and self.dropout_enabled and p and p > 0.0 and self.rng is not None):
      return None
    site = self._site
    self._site += 1
    return (self.rng, float(p), site)

  def _count_batch(self, bn):
    """BatchNorm.num_batches_tracked += 1.  Under a Trainer all the counters are views of one int64 buffer that is
    bumped once per step (training.FlatState.count_batches) instead of one tiny kernel per BatchNorm."""
    if getattr(self, 'batch_counters_fused', False):
      return
    seen = getattr(self, 'bn_seen', None)
    if seen is not None:
      seen.append(bn)
    bn.num_batches_tracked += 1

  def new_arena(self, device, n_floats=6 * 1024 * 1024):
    """One zero-filled fp32 arena per step for all the small accumulators (BatchNorm statistics, SE squeezes, ...):
    a single memset instead of ~500 tiny fill kernels."""
    self._arena = torch.zeros(n_floats, dtype=F32, device=device)
    self._arena_off = 0

  def zeros(self, shape, device):
    n = 1
    for d in shape:
      n *= d
    n4 = (n + 3) & ~3
    arena = getattr(self, '_arena', None)
    if arena is None or arena.device != device or self._arena_off + n4 > arena.numel():
      return torch.zeros(shape, dtype=F32, device=device)
    out = arena[self._arena_off:self._arena_off + n].view(shape)
    self._arena_off += n4
    return out

  def _tap(self, name, t):
    if self.debug_taps is not None:
      self.debug_taps[name] = t

  def conv_bn(self, a, cna, training, *, taps=ops.TAPS_1X1, batch=None, grouped=False, stride=1, act=ACT_NONE, res=None,
              res_bn=None, want_pool=False, a_src=None):
    """ConvNormAct (timm ConvBnAct): conv -> BatchNorm2d -> act, optionally (+ res) before act and per-sample channel
    sums for squeeze-excite.  Training: batch statistics from the conv epilogue, one apply pass.  Eval: everything in
    the conv epilogue.  res_bn = (raw, scale, shift): residual that still needs its own BatchNorm affine.
    grouped: the RegNet 3x3 group conv (stride 1 or 2) on the haloed-tile kernel; everything else is an implicit GEMM."""
    bn = cna.bn
    cout = cna.conv.weight.shape[0]
    if grouped:
      assert ops.gconv3x3_supported(cout, cna.conv.weight.shape[1]), 'group width 24, channels % 72 == 0'
      w = packed(cna.conv.weight, 'gconv_halo')
    else:
      w = packed(cna.conv.weight, 'conv')
    b = a.shape[0] if batch is None else batch
    pool = self.zeros((b, cout), a.device) if want_pool else None
    if training:
      stats = self.zeros((2, cout), a.device)
      if grouped and HALO_UMMA_GCONV and stride == 1:  # EXPERIMENTAL (TFPP_HALO_UMMA_GCONV=1)
        raw = ops.halo_gconv3x3(a, packed(cna.conv.weight, 'gconv_halo_umma'), stats=(stats[0], stats[1]))
      elif grouped:
        raw = ops.gconv3x3(a, w, stride, stats=(stats[0], stats[1]))
      else:
        raw = ops.conv_gemm(a, w, taps=taps, batch=batch, stats=(stats[0], stats[1]))
      count = raw.shape[0] * raw.shape[1] * raw.shape[2]
      scale, shift, mean, invstd = ops.bn_finalize(stats[0], stats[1], bn.weight, bn.bias, bn.running_mean,
                                                   bn.running_var, count, eps=bn.eps, momentum=bn.momentum,
                                                   save=self.tape is not None)
      self._count_batch(bn)
      if res_bn is not None:
        y = ops.scale_shift_act(raw, scale, shift, act, res=res_bn[0], res_scale=res_bn[1], res_shift=res_bn[2],
                                pool_sum=pool)
      else:
        y = ops.scale_shift_act(raw, scale, shift, act, res=res, pool_sum=pool)
      self._save(op='conv_bn', a=a, a_src=a_src, raw=raw, y=y, mean=mean, invstd=invstd, scale=scale, shift=shift,
                 cna=cna, taps=taps,
                 batch=batch, grouped=grouped, stride=stride, act=act, res=res, res_bn=res_bn)
      return (y, pool) if want_pool else y
    scale, shift = packed((bn.weight, bn.bias, bn.running_mean, bn.running_var), 'bn_eval', bn.eps)
    if grouped and HALO_UMMA_GCONV and stride == 1:
      assert res is None
      y = ops.halo_gconv3x3(a, packed(cna.conv.weight, 'gconv_halo_umma'), scale=scale, shift=shift, act=act)
    elif grouped:
      assert res is None
      y = ops.gconv3x3(a, w, stride, scale=scale, shift=shift, act=act)
    else:
      y = ops.conv_gemm(a, w, taps=taps, batch=batch, scale=scale, shift=shift, act=act, res1=res)
    if want_pool:
      ops.scale_shift_act(y, pool_sum=pool, out=y)
      return y, pool
    return y

  def conv_bias(self, a, conv, act=ACT_NONE, taps=None, **kw):
    """nn.Conv2d with bias (+activation) as one implicit-GEMM launch."""
    k = conv.weight.shape[-1]
    taps = taps or (ops.TAPS_3X3 if k == 3 else ops.TAPS_1X1)
    cout, cin = conv.weight.shape[0], conv.weight.shape[1]
    cpad = 8 if cout <= 8 else (16 if cout <= 16 else 32)
    smallc = (a.dtype == BF16 and k == 3 and cout <= 32 and ops.smallc_supported(cin, cpad) and a.shape[1] * a.shape[2] >= 4096 and
              (kw.get('out_layout') == 'nchw' or cout == cpad) and not set(kw) - {'out_layout', 'out_f32'})
    npad = 16 if cout <= 16 else 32
    halo = (smallc and HALO_UMMA and ops.halo_umma_supported(cin, npad) and
            (kw.get('out_layout') == 'nchw' or cout == npad))
    if halo:  # EXPERIMENTAL (TFPP_HALO_UMMA=1): the same layers on tcgen05 (csrc/halo_umma.cu)
      y = ops.halo_conv3x3(a, packed(conv.weight, 'conv_halo_umma', npad), bias=packed(conv.bias, 'f32'), act=act,
                           n_valid=cout, out_nchw_f32=kw.get('out_layout') == 'nchw')
    elif smallc:  # high-resolution, few channels: haloed shared-memory tile kernel (HBM-bound layers)
      y = ops.smallc_conv3x3(a, packed(conv.weight, 'conv_rows_pad', cpad), bias=packed(conv.bias, 'f32'), act=act,
                             n_valid=cout, out_nchw_f32=kw.get('out_layout') == 'nchw')
    else:
      y = ops.conv_gemm(a, packed(conv.weight, 'conv'), taps=taps, shift=packed(conv.bias, 'f32'), act=act, **kw)
    self._save(op='conv_bias', a=a, y=y, conv=conv, act=act, taps=taps, kw=kw, smallc=smallc, halo=halo)
    return y

  # ------------------------------------------------------------------------------------------------ RegNet
  def regnet_block(self, x, blk, training):
    """timm regnet.Bottleneck.forward (oracle/regnety.py): conv1 -> conv2 (grouped, stride) -> SE -> conv3 -> +shortcut
    -> ReLU."""
    b, h, w, _ = x.shape
    s = blk.stride
    a1 = self.conv_bn(x, blk.conv1, training, act=ACT_RELU)
    a2, pool = self.conv_bn(a1, blk.conv2, training, grouped=True, stride=s, act=ACT_RELU, want_pool=True)
    ho, wo = h // s, w // s
    se = blk.se
    hidden = None
    if self.tape is not None:
      gate, hidden = ops.se_gate(pool, ho * wo, se.fc1.weight, se.fc1.bias, se.fc2.weight, se.fc2.bias,
                                 want_hidden=True)
    else:
      gate = ops.se_gate(pool, ho * wo, se.fc1.weight, se.fc1.bias, se.fc2.weight, se.fc2.bias)
    a2s = ops.channel_scale(a2, gate)
    self._save(op='se', a2=a2, a2s=a2s, pool=pool, gate=gate, hidden=hidden, se=se, hw=ho * wo)
    if blk.downsample is not None:
      xp = ops.parity_split(x) if s == 2 else x
      ds = blk.downsample
      if training:
        # raw downsample conv + its batch statistics; its BatchNorm affine is applied inside conv3's apply pass
        wd = packed(ds.conv.weight, 'conv')
        stats = self.zeros((2, wd.shape[0]), x.device)
        raw_d = ops.conv_gemm(xp, wd, batch=b, stats=(stats[0], stats[1]))
        count = b * ho * wo
        sd, td, mean_d, invstd_d = ops.bn_finalize(stats[0], stats[1], ds.bn.weight, ds.bn.bias, ds.bn.running_mean,
                                                   ds.bn.running_var, count, eps=ds.bn.eps, momentum=ds.bn.momentum,
                                                   save=self.tape is not None)
        self._count_batch(ds.bn)
        self._save(op='downsample', a=xp, x_src=x, stride=s, raw=raw_d, mean=mean_d, invstd=invstd_d, cna=ds, batch=b)
        return self.conv_bn(a2s, blk.conv3, training, act=ACT_RELU, res_bn=(raw_d, sd, td))
      shortcut = self.conv_bn(xp, ds, training, batch=b)
      return self.conv_bn(a2s, blk.conv3, training, act=ACT_RELU, res=shortcut)
    return self.conv_bn(a2s, blk.conv3, training, act=ACT_RELU, res=x)

  def regnet_stage(self, x, stage, training):
    for blk in stage:
      x = self.regnet_block(x, blk, training)
    return x

  def stem(self, x, cna, training, normalize):
    """timm stem ConvNormAct (3x3, stride 2) fused with normalize_imagenet (transfuser_utils.py:542-551)."""
    dev = x.device
    in_scale = in_shift = None
    if normalize:
      in_scale = self._const('im_scale', lambda: torch.tensor([1.0 / (255.0 * s) for s in IMAGENET_STD]), dev)
      in_shift = self._const('im_shift', lambda: torch.tensor([-m / s for m, s in zip(IMAGENET_MEAN, IMAGENET_STD)]),
                             dev)
    w = packed(cna.conv.weight, 'f32')
    bn = cna.bn
    if training:
      stats = self.zeros((2, 32), dev)
      raw = ops.stem_conv(x, w, in_scale, in_shift, stats=(stats[0], stats[1]))
      count = raw.shape[0] * raw.shape[1] * raw.shape[2]
      scale, shift, mean, invstd = ops.bn_finalize(stats[0], stats[1], bn.weight, bn.bias, bn.running_mean,
                                                   bn.running_var, count, eps=bn.eps, momentum=bn.momentum,
                                                   save=self.tape is not None)
      self._count_batch(bn)
      y = ops.scale_shift_act(raw, scale, shift, ACT_RELU)
      self._save(op='stem', x=x, raw=raw, y=y, mean=mean, invstd=invstd, scale=scale, shift=shift, cna=cna,
                 in_scale=in_scale,
                 in_shift=in_shift)
      return y
    scale, shift = packed((bn.weight, bn.bias, bn.running_mean, bn.running_var), 'bn_eval', bn.eps)
    return ops.stem_conv(x, w, in_scale, in_shift, scale=scale, shift=shift, act=ACT_RELU)

  # ------------------------------------------------------------------------------------------------ fusion GPT
  def fuse(self, img, lid, i, training):
    """TransfuserBackbone.fuse_features + GPT.forward (transfuser.py:222-257,301-339) for scale i."""
    bb, cfg = self.bb, self.cfg
    gpt = bb.transformers[i]
    b, hi, wi, c = img.shape
    _, hl, wl, cl = lid.shape
    ph_i, pw_i, ph_l, pw_l = cfg.img_vert_anchors, cfg.img_horz_anchors, cfg.lidar_vert_anchors, cfg.lidar_horz_anchors
    n_img, n_lid = ph_i * pw_i, ph_l * pw_l
    t = n_img + n_lid
    dev = img.device
    pos = packed(gpt.pos_emb, 'f32').view(t, c)
    x = torch.empty((b, t, c), dtype=F32, device=dev)  # fp32 residual stream
    ops.avgpool_tokens(img, x, ph_i, pw_i, 0, pos_emb=pos)
    lid_pool = torch.empty((b, n_lid, cl), dtype=ops.act_dtype(), device=dev)
    ops.avgpool_tokens(lid, lid_pool, ph_l, pw_l, 0)
    l2i = bb.lidar_channel_to_img[i]
    ops.linear(lid_pool.view(b * n_lid, cl), packed(l2i.weight, 'linear'), bias=packed(l2i.bias, 'f32'),
               out=x.view(-1)[n_img * c:], row_map=(n_lid, t), res2=pos[n_img:], res2_strides=(0, 0, c, 1))
    x = x.view(b * t, c)
    d_embd = self.drop(cfg.embd_pdrop, training)  # GPT.drop on pos_emb + tokens (transfuser.py:325)
    ops.dropout_(x, d_embd)
    self._save(op='tokenise', img=img, lid=lid, lid_pool=lid_pool, i=i, x0=x, b=b, t=t, c=c, cl=cl, drop=d_embd)
    heads = cfg.n_head
    for blk in gpt.blocks:
      at = blk.attn
      h, _, mean1, rstd1 = ops.layernorm(x, blk.ln1.weight, blk.ln1.bias, save=self.tape is not None)
      wqkv = packed((at.query.weight, at.key.weight, at.value.weight), 'cat_linear')
      bqkv = packed((at.query.bias, at.key.bias, at.value.bias), 'cat_f32')
      qkv = ops.linear(h, wqkv, bias=bqkv)
      d_attn = self.drop(cfg.attn_pdrop, training)    # attn_drop on the probabilities (transfuser.py:374)
      y = ops.fusion_attn(qkv, b, t, c, heads, drop=d_attn)
      d_proj = self.drop(cfg.resid_pdrop, training)   # resid_drop (transfuser.py:379): x + drop(proj(y))
      x1 = ops.linear(y, packed(at.proj.weight, 'linear'), bias=packed(at.proj.bias, 'f32'), res=x, out_f32=True,
                      drop=d_proj)
      h2, _, mean2, rstd2 = ops.layernorm(x1, blk.ln2.weight, blk.ln2.bias, save=self.tape is not None)
      m = ops.linear(h2, packed(blk.mlp[0].weight, 'linear'), bias=packed(blk.mlp[0].bias, 'f32'), act=ACT_RELU)
      d_mlp = self.drop(cfg.resid_pdrop, training)    # nn.Dropout closing the MLP (transfuser.py:395)
      x2 = ops.linear(m, packed(blk.mlp[2].weight, 'linear'), bias=packed(blk.mlp[2].bias, 'f32'), res=x1, out_f32=True,
                      drop=d_mlp)
      self._save(op='gpt_block', x=x, h=h, qkv=qkv, y=y, x1=x1, h2=h2, m=m, x2=x2, blk=blk, mean1=mean1, rstd1=rstd1,
                 mean2=mean2, rstd2=rstd2, b=b, t=t, c=c, heads=heads, drops=(d_attn, d_proj, d_mlp))
      x = x2
    xf, _, meanf, rstdf = ops.layernorm(x, gpt.ln_f.weight, gpt.ln_f.bias, save=self.tape is not None)
    # image tokens: bilinear up-sample straight out of the token matrix + residual add (transfuser.py:239-242,254)
    img_out = ops.bilinear(xf, b, ph_i, pw_i, hi, wi, c, src_batch_stride=t * c, src_row_stride=c, add=img)
    # LiDAR tokens: 1x1 conv back to the LiDAR width on the 64-row slab, then up-sample + add (transfuser.py:237,250-255)
    i2l = bb.img_channel_to_lidar[i]
    lid_tok = torch.empty((b * n_lid, cl), dtype=ops.act_dtype(), device=dev)
    ops.conv_gemm(xf.view(-1)[n_img * c:], packed(i2l.weight, 'linear').view(cl, 1, c), a_shape=(b, 1, n_lid, c),
                  a_batch_stride=t * c, shift=packed(i2l.bias, 'f32'), out=lid_tok, out_strides=(n_lid * cl, 0, cl, 1))
    lid_out = ops.bilinear(lid_tok, b, ph_l, pw_l, hl, wl, cl, add=lid)
    self._save(op='untokenise', x=x, xf=xf, lid_tok=lid_tok, i=i, meanf=meanf, rstdf=rstdf, b=b, t=t, c=c, cl=cl,
               img=img, lid=lid, img_out=img_out, lid_out=lid_out)
    return img_out, lid_out

  # ------------------------------------------------------------------------------------------------ backbone
  def backbone_forward(self, image, lidar, training):
    """TransfuserBackbone.forward (transfuser.py:139-205). image (B,3,H,W) f32 0..255, lidar (B,C,256,256) f32.
    Returns NHWC bf16 (bev features (B,64,64,64), fused LiDAR features (B,8,8,1512), image grid (B,8,32,1512))."""
    bb, cfg = self.bb, self.cfg
    if cfg.backbone == 'bev_encoder':
      return self.bev_backbone_forward(image, lidar, training)
    if not (image.is_cuda and lidar.is_cuda):
      raise RuntimeError('carla_garage_b200 runs on CUDA tensors only (no CPU fallback)')
    if training:
      # tfpp_bn_finalize / tfpp_extra_sensor_token update the running statistics through raw pointers (no version
      # counter moves): folded eval-mode BatchNorm affines cached before this forward are stale after it
      PARAM_EPOCH[0] += 1
      if self.dropout_enabled:
        self.begin_dropout_step(image.device)  # on the main stream, before the LiDAR branch forks
    image = image.float().contiguous()
    lidar = lidar.float().contiguous()
    # Training: between two fusion points the LiDAR branch (a quarter of the image branch's pixels: kernels that cannot
    # fill 148 SMs) runs on a second stream next to the image branch.  Only with a tape: every tensor that crosses
    # streams is then kept alive by a tape record until the step ends, which is what makes this safe with the caching
    # allocator; the tape records are tagged so that the backward pass forks / joins the same way.
    two = self.tape is not None and os.environ.get('TFPP_NO_OVERLAP', '0') != '1'
    main = torch.cuda.current_stream()
    side = self.side_stream(image.device, 'lidar') if two else None

    def on_side(fn):
      if not two:
        return fn()
      side.wait_stream(main)
      first = len(self.tape)
      with torch.cuda.stream(side):
        out = fn()
      for r in self.tape[first:]:
        r['side'] = 'lidar'
      return out

    def joined(fn):
      if not two:
        return fn()
      main.wait_stream(side)
      first = len(self.tape)
      out = fn()
      for r in self.tape[first:]:
        r['needs_side'] = 'lidar'
      return out

    lid = on_side(lambda: self.stem(lidar, bb.lidar_encoder['stem'], training, False))
    img = self.stem(image, bb.image_encoder['stem'], training, cfg.normalize_imagenet)
    self._tap('img_stem', img)
    self._tap('lid_stem', lid)
    for i in range(4):
      lid = on_side(lambda lid=lid: self.regnet_stage(lid, bb.lidar_encoder[f's{i + 1}'], training))
      img = self.regnet_stage(img, bb.image_encoder[f's{i + 1}'], training)
      self._tap(f'img_s{i + 1}_pre', img)
      self._tap(f'lid_s{i + 1}_pre', lid)
      img, lid = joined(lambda img=img, lid=lid: self.fuse(img, lid, i, training))
      self._tap(f'img_s{i + 1}', img)
      self._tap(f'lid_s{i + 1}', lid)
    feats = None
    if cfg.detect_boxes or cfg.use_bev_semantic:
      # top_down (transfuser.py:131-137)
      b = lid.shape[0]
      p5 = self.conv_bias(lid, bb.c5_conv, ACT_RELU)
      up = cfg.bev_upsample_factor
      p5u = ops.bilinear(p5, b, p5.shape[1], p5.shape[2], p5.shape[1] * up, p5.shape[2] * up, p5.shape[3])
      self._save(op='bilinear', src=p5, out=p5u)
      p4 = self.conv_bias(p5u, bb.up_conv5, ACT_RELU)
      th = cfg.lidar_resolution_height // cfg.bev_down_sample_factor
      tw = cfg.lidar_resolution_width // cfg.bev_down_sample_factor
      p4u = ops.bilinear(p4, b, p4.shape[1], p4.shape[2], th, tw, p4.shape[3])
      self._save(op='bilinear', src=p4, out=p4u)
      feats = self.conv_bias(p4u, bb.up_conv4, ACT_RELU)
    grid = img if (cfg.use_semantic or cfg.use_depth) else None
    self._tap('bev_feature_grid', feats)
    self._tap('image_feature_grid', grid)
    if not cfg.transformer_decoder_join:
      # transfuser.py:188-197: global average pools, lidar_to_img_features_end, sum -> (B, num_features)
      b, c, cl = img.shape[0], img.shape[3], lid.shape[3]
      img_pool = torch.empty((b, 1, c), dtype=ops.act_dtype(), device=img.device)
      ops.avgpool_tokens(img, img_pool, 1, 1, 0)
      lid_pool = torch.empty((b, 1, cl), dtype=ops.act_dtype(), device=img.device)
      ops.avgpool_tokens(lid, lid_pool, 1, 1, 0)
      end = bb.lidar_to_img_features_end
      fused = ops.linear(lid_pool.view(b, cl), packed(end.weight, 'linear'), bias=packed(end.bias, 'f32'),
                         res=img_pool.view(b, c))
      self._save(op='global_fuse', img=img, lid=lid, lid_pool=lid_pool, fused=fused, b=b, c=c, cl=cl)
      self._tap('fused_features', fused)
      return feats, fused, grid
    self._tap('fused_features', lid)
    return feats, lid, grid

  # ------------------------------------------------------------------------------------------------ bev_encoder
  def conv_in(self, inputs, conv, norm, act, out=None):
    """Conv2d(3x3, bias=False) over the channel concatenation of ``inputs`` -> nn.InstanceNorm2d -> act
    (bev_encoder.py:126-137,253-262).  The concatenation is never built: conv(cat(a, b)) = conv_a(a) + conv_b(b) with
    the weight sliced along Cin, partial sums in fp32.  ``out``: wider NHWC tensor whose first Cout channels receive y."""
    raw, k0, slices = None, 0, []
    for i, a in enumerate(inputs):
      k1 = k0 + a.shape[3]
      raw = ops.conv_gemm(a, packed(conv.weight, 'conv_cin', k0, k1), taps=ops.TAPS_3X3, res1=raw,
                          out_f32=i + 1 < len(inputs))
      slices.append((a, k0, k1))
      k0 = k1
    ps = out.shape[3] if out is not None else None
    y, mean, invstd = ops.instnorm(raw, act, norm.eps, out=out, out_pix_stride=ps, zeros=self.zeros,
                                   save=self.tape is not None)
    self._save(op='conv_in', slices=slices, raw=raw, y=y, mean=mean, invstd=invstd, conv=conv, act=act, y_pix_stride=ps)
    return y

  def lift_tables(self, device, img_h, img_w):
    """tfpp_bev_lift tables folded from the module's grid / normaliser / mask parameters (nn.bev_encoder.lift_tables)."""
    bb = self.bb
    ps = (bb.grid, bb.bev_projection_normalizer, bb.valid_bev_pixels)
    ver = tuple(p._version for p in ps) + tuple(p.data_ptr() for p in ps) + (img_h, img_w, str(device))  # pylint: disable=protected-access
    hit = self._consts.get('lift_tables')
    if hit is None or hit[0] != ver:
      from .nn.bev_encoder import lift_tables  # pylint: disable=import-outside-toplevel
      hit = (ver, tuple(t.to(device) for t in lift_tables(*ps, img_h, img_w)))
      self._consts['lift_tables'] = hit
    return hit[1]

  def bev_stem(self, cat, cna, training):
    """Stem ConvNormAct of the BEV RegNet (3x3, stride 2) over [compressed camera features | LiDAR | zero padding]:
    one implicit GEMM over the four parity planes (ops.taps_3x3_stride2) with the BatchNorm statistics in its epilogue."""
    b, cpad = cat.shape[0], cat.shape[3]
    planes = ops.parity_split(cat)
    w = packed(cna.conv.weight, 'conv_cin_pad', cpad)
    taps = ops.taps_3x3_stride2(b)
    bn = cna.bn
    if training:
      stats = self.zeros((2, w.shape[0]), cat.device)
      raw = ops.conv_gemm(planes, w, taps=taps, batch=b, stats=(stats[0], stats[1]))
      count = raw.shape[0] * raw.shape[1] * raw.shape[2]
      scale, shift, mean, invstd = ops.bn_finalize(stats[0], stats[1], bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                                   count, eps=bn.eps, momentum=bn.momentum, save=self.tape is not None)
      self._count_batch(bn)
      y = ops.scale_shift_act(raw, scale, shift, ACT_RELU)
      self._save(op='bev_stem', cat=cat, planes=planes, raw=raw, y=y, mean=mean, invstd=invstd, scale=scale, shift=shift,
                 cna=cna)
      return y
    scale, shift = packed((bn.weight, bn.bias, bn.running_mean, bn.running_var), 'bn_eval', bn.eps)
    return ops.conv_gemm(planes, w, taps=taps, batch=b, scale=scale, shift=shift, act=ACT_RELU)

  def bev_backbone_forward(self, image, lidar, training):
    """BevEncoder.forward (bev_encoder.py:146-233).  image (B,3,H,W) f32 0..255, lidar (B,C,256,256) f32.  Returns NHWC
    (bev feature grid (B,64,64,64), fused BEV features (B,16,16,576), perspective image features (B,32,128,32))."""
    bb, cfg = self.bb, self.cfg
    if not (image.is_cuda and lidar.is_cuda):
      raise RuntimeError('carla_garage_b200 runs on CUDA tensors only (no CPU fallback)')
    if not cfg.transformer_decoder_join:
      raise NotImplementedError('bev_encoder with the global-pool MLP join is not built')
    if training:
      PARAM_EPOCH[0] += 1
      if self.dropout_enabled:
        self.begin_dropout_step(image.device)
    image = image.float().contiguous()
    lidar = lidar.float().contiguous()
    b = image.shape[0]
    enc = bb.image_encoder
    x = self.stem(image, enc['stem'], training, cfg.normalize_imagenet)
    x = self.regnet_stage(x, enc['s1'], training)
    x2 = self.regnet_stage(x, enc['s2'], training)
    x3 = self.regnet_stage(x2, enc['s3'], training)
    self._tap('img_s2', x2)
    self._tap('img_s3', x3)
    # UpsamplingConcat (bev_encoder.py:264-272) + depth_layer
    up = ops.bilinear(x3, b, x3.shape[1], x3.shape[2], x2.shape[1], x2.shape[2], x3.shape[3])
    self._save(op='bilinear', src=x3, out=up)
    ul = bb.upsampling_layer.conv
    u = self.conv_in([x2, up], ul[0], ul[1], ACT_RELU)
    u = self.conv_in([u], ul[3], ul[4], ACT_RELU)
    self._tap('upsampled', u)
    feat = self.conv_bias(u, bb.depth_layer)
    self._tap('image_features', feat)
    # lift to BEV (bev_encoder.py:179-199) and compress (:126-137) straight into the stem's input tensor
    depth, width = bb.grid.shape[1], bb.grid.shape[2]
    tables = self.lift_tables(image.device, feat.shape[1], feat.shape[2])
    bev = ops.bev_lift(feat, tables, depth, width)
    self._save(op='bev_lift', img=feat, out=bev, tables=tables)
    self._tap('bev_lift', bev)
    cl, cb = lidar.shape[1], bev.shape[3]
    cpad = (cb + cl + 7) // 8 * 8
    cat = torch.zeros((b, width, depth, cpad), dtype=ops.act_dtype(), device=image.device)
    comp = bb.bev_compressor
    self.conv_in([bev], comp[0], comp[1], ACT_GELU, out=cat)
    cat[..., cb:cb + cl].copy_(lidar.permute(0, 2, 3, 1))   # torch.cat((bev_features, lidar_features), dim=1)
    self._tap('bev_cat', cat)
    f = self.bev_stem(cat, bb.bev_encoder['stem'], training)
    for i in (1, 2, 3):
      f = self.regnet_stage(f, bb.bev_encoder[f's{i}'], training)
      self._tap(f'bev_s{i}', f)
    feats = None
    if cfg.detect_boxes or cfg.use_bev_semantic:  # top_down (bev_encoder.py:139-144)
      p5 = self.conv_bias(f, bb.c5_conv, ACT_RELU)
      up_f = cfg.bev_upsample_factor
      p5u = ops.bilinear(p5, b, p5.shape[1], p5.shape[2], p5.shape[1] * up_f, p5.shape[2] * up_f, p5.shape[3])
      self._save(op='bilinear', src=p5, out=p5u)
      p4 = self.conv_bias(p5u, bb.up_conv5, ACT_RELU)
      th = cfg.lidar_resolution_height // cfg.bev_down_sample_factor
      tw = cfg.lidar_resolution_width // cfg.bev_down_sample_factor
      p4u = ops.bilinear(p4, b, p4.shape[1], p4.shape[2], th, tw, p4.shape[3])
      self._save(op='bilinear', src=p4, out=p4u)
      feats = self.conv_bias(p4u, bb.up_conv4, ACT_RELU)
    self._tap('bev_feature_grid', feats)
    self._tap('fused_features', f)
    return feats, f, feat

  # ------------------------------------------------------------------------------------------------ heads
  def perspective_decoder(self, dec, grid, act_last=ACT_NONE):
    """t_u.PerspectiveDecoder.forward (transfuser_utils.py:697-704); last conv writes NCHW f32."""
    b = grid.shape[0]
    x = self.conv_bias(grid, dec.deconv1[0], ACT_RELU)
    x = self.conv_bias(x, dec.deconv1[2], ACT_RELU)
    s0 = dec.scale_factor_0
    xu = ops.bilinear(x, b, x.shape[1], x.shape[2], x.shape[1] * s0, x.shape[2] * s0, x.shape[3])
    self._save(op='bilinear', src=x, out=xu)
    x = self.conv_bias(xu, dec.deconv2[0], ACT_RELU)
    x = self.conv_bias(x, dec.deconv2[2], ACT_RELU)
    s1 = dec.scale_factor_1
    xu = ops.bilinear(x, b, x.shape[1], x.shape[2], x.shape[1] * s1, x.shape[2] * s1, x.shape[3])
    self._save(op='bilinear', src=x, out=xu)
    x = self.conv_bias(xu, dec.deconv3[0], ACT_RELU)
    return self.conv_bias(x, dec.deconv3[2], act_last, out_layout='nchw', out_f32=True)

  def center_head_forward(self, feat):
    """LidarCenterNetHead.forward (center_net.py:49-75): the five 3x3 convs as one N=320 implicit GEMM, the five 1x1
    convs as one block-diagonal GEMM writing a (B,21,64,64) NCHW f32 buffer (sigmoid on the 4 heat-map channels)."""
    head = self.head
    names = head.head_names()
    convs0 = tuple(getattr(head, n)[0] for n in names)
    convs1 = tuple(getattr(head, n)[2] for n in names)
    w0 = packed(tuple(c.weight for c in convs0), 'cat_conv')
    b0 = packed(tuple(c.bias for c in convs0), 'cat_f32')
    h = ops.conv_gemm(feat, w0, taps=ops.TAPS_3X3, shift=b0, act=ACT_RELU)
    w1 = packed(tuple(c.weight for c in convs1), 'blockdiag_1x1')
    b1 = packed(tuple(c.bias for c in convs1), 'cat_f32')
    ncls = convs1[0].weight.shape[0]
    out = ops.conv_gemm(h, w1, shift=b1, act=ACT_SIGMOID, act_n_limit=ncls, out_layout='nchw', out_f32=True)
    self._save(op='center_head', feat=feat, h=h, out=out, convs0=convs0, convs1=convs1)
    sizes = [c.weight.shape[0] for c in convs1]
    views, o = [], 0
    for s in sizes:
      views.append(out[:, o:o + s])
      o += s
    return (views[0], views[1], views[2], views[3], views[4], None, None)

  def planner(self, fused, target_point, ego_vel, command, training):
    """model.py:299-358: memory tokens, then one pass of the 6-layer decoder per query set — ``wp_query`` -> wp_decoder
    (use_wp_gru, model.py:325-337) and ``checkpoint_query`` -> checkpoint_decoder + target-speed MLP
    (use_controller_input_prediction, model.py:338-358).  Returns (pred_checkpoint, pred_target_speed, pred_wp)."""
    m, cfg = self.m, self.cfg
    b, fh, fw, cf = fused.shape
    dev = fused.device
    d = cfg.gru_input_size
    n_pix = fh * fw
    n_mem = n_pix + 1
    mem = torch.empty((b, n_mem, d), dtype=ops.act_dtype(), device=dev)
    posenc = self._const(f'posenc{fh}x{fw}', lambda: m.encoder_pos_encoding.table(fh, fw), dev)
    cc = m.change_channel
    ops.linear(fused.view(b * n_pix, cf), packed(cc.weight, 'linear'), bias=packed(cc.bias, 'f32'), out=mem,
               row_map=(n_pix, n_mem), res2=posenc, res2_strides=(0, 0, d, 1))
    vn = m.velocity_normalization
    ese = m.extra_sensor_encoder
    vmean, vvar = (0.0, 1.0) if training else packed((vn.running_mean, vn.running_var), 'host_floats')
    ops.extra_sensor_token(ego_vel.float().contiguous(), command.float().contiguous(), vmean, vvar, training,
                           vn.running_mean if training else None,
                           vn.running_var if training else None, ese[0].weight, ese[0].bias, ese[2].weight, ese[2].bias,
                           packed(m.extra_sensor_pos_embed, 'f32'), mem if mem.dtype == BF16 else None,
                           mem if mem.dtype == F32 else None, n_mem, n_pix)
    if training:
      self._count_batch(vn)
    memf = mem.view(b * n_mem, d)
    kvs = []
    for l in m.join.layers:  # K/V projections of the (layer- and query-independent) memory: one small GEMM per layer
      kvs.append(ops.linear(memf, packed(l.multihead_attn.in_proj_weight, 'rows', d, 3 * d),
                            bias=packed(l.multihead_attn.in_proj_bias, 'rows_f32', d, 3 * d)))
    self._save(op='planner_mem', fused=fused, mem=mem, kvs=kvs, b=b, n_pix=n_pix, n_mem=n_mem, d=d,
               ego_vel=ego_vel, command=command, training=training, posenc=posenc)
    tp = target_point.float().contiguous()
    pred_wp = pred_cp = pred_ts = None
    if getattr(cfg, 'use_wp_gru', False):
      joined = self._decode(m.wp_query, kvs, b, n_mem, training)
      pred_wp = self._gru_head(joined, m.wp_decoder, None, tp, 'planner_wp')[0]
    if cfg.use_controller_input_prediction:
      joined = self._decode(m.checkpoint_query, kvs, b, n_mem, training)
      pred_cp, pred_ts = self._gru_head(joined, m.checkpoint_decoder, m.target_speed_network, tp, 'planner')
      self._tap('joined', joined[1].view(b, -1, d))
    return pred_cp, pred_ts, pred_wp

  def planner_mlp(self, fused, target_point, ego_vel, command, training):
    """model.py:306-322,359-376 with transformer_decoder_join = False (the original TransFuser planner): extra-sensor
    embedding ++ globally pooled features -> 3-layer MLP join -> autoregressive GRUCell heads
    (GRUWaypointsPredictorTransFuser) + the target-speed MLP on the first gru_hidden_size features.
    Returns (pred_checkpoint, pred_target_speed, pred_wp)."""
    m, cfg = self.m, self.cfg
    b, c = fused.shape
    dev = fused.device
    hs, e = cfg.gru_hidden_size, cfg.extra_sensor_channels
    vn, ese = m.velocity_normalization, m.extra_sensor_encoder
    es = torch.empty((b, 1, e), dtype=ops.act_dtype(), device=dev)
    zero_pos = self._const(f'zeros{e}', lambda: torch.zeros(e), dev)
    vmean, vvar = (0.0, 1.0) if training else packed((vn.running_mean, vn.running_var), 'host_floats')
    ops.extra_sensor_token(ego_vel.float().contiguous(), command.float().contiguous(), vmean, vvar, training,
                           vn.running_mean if training else None, vn.running_var if training else None, ese[0].weight,
                           ese[0].bias, ese[2].weight, ese[2].bias, zero_pos, es if es.dtype == BF16 else None,
                           es if es.dtype == F32 else None, 1, 0)
    if training:
      self._count_batch(vn)
    j0, j1, j2 = m.join[0], m.join[2], m.join[4]
    # Linear over the concatenation [fused | extra sensors] (model.py:322) = two GEMMs on the column slices of its weight
    h1a = ops.linear(fused, packed(j0.weight, 'cols', 0, c), out_f32=True)
    h1 = ops.linear(es.view(b, e), packed(j0.weight, 'cols', c, c + e), bias=packed(j0.bias, 'f32'), res=h1a, act=ACT_RELU)
    h2 = ops.linear(h1, packed(j1.weight, 'linear'), bias=packed(j1.bias, 'f32'), act=ACT_RELU)
    joined = ops.linear(h2, packed(j2.weight, 'linear'), bias=packed(j2.bias, 'f32'), act=ACT_RELU, out_f32=True)
    self._tap('joined', joined)
    self._save(op='mlp_join', fused=fused, es=es, h1=h1, h2=h2, joined=joined, b=b, c=c, e=e, ego_vel=ego_vel,
               command=command, training=training)
    tp = target_point.float().contiguous() if cfg.use_tp else None
    pred_wp = pred_cp = pred_ts = None
    if getattr(cfg, 'use_wp_gru', False):
      pred_wp = self._gru_cell_head(joined, m.wp_decoder, None, tp, 'planner_wp')[0]
    if cfg.use_controller_input_prediction:
      pred_cp, pred_ts = self._gru_cell_head(joined, m.checkpoint_decoder, m.target_speed_network, tp, 'planner')
    return pred_cp, pred_ts, pred_wp

  def _gru_cell_head(self, joined, cd, tsn, target_point, seed_key):
    """GRUWaypointsPredictorTransFuser.forward (model.py:886-913) [+ target_speed_network, model.py:376]."""
    cell = cd.wp_decoder
    ts_w = (tsn[0].weight, tsn[0].bias, tsn[2].weight, tsn[2].bias) if tsn is not None else (None,) * 4
    wp, ts, h_all = ops.gru_cell_head(joined, target_point, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh,
                                      cd.output.weight, cd.output.bias, *ts_w, steps=cd.prediction_len,
                                      hidden=cd.hidden_size, learn_origin=bool(self.cfg.learn_origin),
                                      want_h=self.tape is not None)
    self._save(op='gru_cell_head', joined=joined, wp=wp, h_all=h_all, target_point=target_point, cd=cd, tsn=tsn,
               seed_key=seed_key)
    return wp, ts

  def _decode(self, query, kvs, b, n_mem, training):
    """nn.TransformerDecoder (post-norm layers + final LayerNorm, model.py:137-143,352) over one learned query set.
    Returns (pre-norm x, joined f32, LayerNorm statistics)."""
    m, cfg = self.m, self.cfg
    d = cfg.gru_input_size
    heads = cfg.num_decoder_heads
    hd = d // heads
    nq = query.shape[1]
    x, xb = packed(query, 'repeat_rows', b)
    self._save(op='planner_queries', x0=x, query=query, b=b)
    for li, l in enumerate(m.join.layers):
      act = ACT_RELU if l.activation is torch.nn.functional.relu else ACT_GELU
      # nn.TransformerDecoderLayer(dropout=0.1) (model.py:137-140): probabilities of both attentions, dropout1/2/3 on
      # the sub-layer outputs, `dropout` after the feed-forward activation — numbered in the order torch applies them
      dp = (self.drop(l.self_attn.dropout, training), self.drop(l.dropout1.p, training),
            self.drop(l.multihead_attn.dropout, training), self.drop(l.dropout2.p, training),
            self.drop(l.dropout.p, training), self.drop(l.dropout3.p, training))
      if dp[4] is not None and act != ACT_RELU:
        raise NotImplementedError('feed-forward dropout is fused with the ReLU mask; GELU + dropout is not built')
      qkv = ops.linear(xb, packed(l.self_attn.in_proj_weight, 'linear'), bias=packed(l.self_attn.in_proj_bias, 'f32'))
      sa = ops.small_mha(qkv, qkv, qkv, b, heads, nq, nq, hd, (nq * 3 * d, 3 * d), (nq * 3 * d, 3 * d),
                         (nq * 3 * d, 3 * d), k_off=d, v_off=2 * d, drop=dp[0])
      t1 = ops.linear(sa, packed(l.self_attn.out_proj.weight, 'linear'), bias=packed(l.self_attn.out_proj.bias, 'f32'),
                      res=x, out_f32=True, drop=dp[1])
      x1b, x1, m1, r1 = ops.layernorm(t1, l.norm1.weight, l.norm1.bias, want_f32=True, eps=l.norm1.eps,
                                      save=self.tape is not None)
      q2 = ops.linear(x1b, packed(l.multihead_attn.in_proj_weight, 'rows', 0, d),
                      bias=packed(l.multihead_attn.in_proj_bias, 'rows_f32', 0, d))
      kv = kvs[li]
      ca = ops.small_mha(q2, kv, kv, b, heads, nq, n_mem, hd, (nq * d, d), (n_mem * 2 * d, 2 * d), (n_mem * 2 * d, 2 * d),
                         v_off=d, drop=dp[2])
      t2 = ops.linear(ca, packed(l.multihead_attn.out_proj.weight, 'linear'),
                      bias=packed(l.multihead_attn.out_proj.bias, 'f32'), res=x1, out_f32=True, drop=dp[3])
      x2b, x2, m2, r2 = ops.layernorm(t2, l.norm2.weight, l.norm2.bias, want_f32=True, eps=l.norm2.eps,
                                      save=self.tape is not None)
      ff = ops.linear(x2b, packed(l.linear1.weight, 'linear'), bias=packed(l.linear1.bias, 'f32'), act=act, drop=dp[4])
      t3 = ops.linear(ff, packed(l.linear2.weight, 'linear'), bias=packed(l.linear2.bias, 'f32'), res=x2, out_f32=True,
                      drop=dp[5])
      x3b, x3, m3, r3 = ops.layernorm(t3, l.norm3.weight, l.norm3.bias, want_f32=True, eps=l.norm3.eps,
                                      save=self.tape is not None)
      self._save(op='dec_layer', li=li, layer=l, x_in=x, xb=xb, qkv=qkv, sa=sa, t1=t1, x1=x1, x1b=x1b, q2=q2, kv=kv, ca=ca,
                 t2=t2, x2=x2, x2b=x2b, ff=ff, t3=t3, x3=x3, stats=(m1, r1, m2, r2, m3, r3), act=act, b=b, nq=nq,
                 n_mem=n_mem, d=d, heads=heads, hd=hd, drops=dp)
      x, xb = x3, x3b
    _, joined, mj, rj = ops.layernorm(x, m.join.norm.weight, m.join.norm.bias, want_bf16=False, want_f32=True,
                                      eps=m.join.norm.eps, save=self.tape is not None)
    return x, joined, (mj, rj)

  def _gru_head(self, dec_out, cd, tsn, target_point, seed_key):
    """GRUWaypointsPredictorInterFuser (model.py:839-867) [+ target_speed_network on the last query, model.py:358]."""
    x, joined, stats = dec_out
    d = joined.shape[1]
    b = target_point.shape[0]
    nq = joined.shape[0] // b
    ts_w = (tsn[0].weight, tsn[0].bias, tsn[2].weight, tsn[2].bias) if tsn is not None else (None,) * 4
    res = ops.planner_head(joined.view(b, nq, d), target_point, cd.encoder.weight, cd.encoder.bias, cd.gru.weight_ih_l0,
                           cd.gru.weight_hh_l0, cd.gru.bias_ih_l0, cd.gru.bias_hh_l0, cd.decoder.weight, cd.decoder.bias,
                           *ts_w, want_h=self.tape is not None)
    self._save(op='planner_head', x=x, joined=joined, stats=stats, res=res, target_point=target_point, b=b, nq=nq, d=d,
               cd=cd, tsn=tsn, seed_key=seed_key)
    return res[0], res[1]

  # ------------------------------------------------------------------------------------------------ full model
  def forward(self, rgb, lidar_bev, target_point, ego_vel, command, training=False):
    """LidarCenterNet.forward (model.py:279-392)."""
    m, cfg = self.m, self.cfg
    self.new_arena(rgb.device)
    feats, fused, grid = self.backbone_forward(rgb, lidar_bev, training)
    # the planner (≈100 tiny launches on a 65-token memory) and the dense heads are independent consumers of the
    # backbone: the planner goes to a side stream and overlaps with the high-resolution decoders (fork / join, also
    # under CUDA-graph capture); its tape records are tagged so the backward pass can do the same
    tp, ev_, cmd = target_point.to(rgb.device), ego_vel.to(rgb.device), command.to(rgb.device)
    overlap = rgb.is_cuda and os.environ.get('TFPP_NO_OVERLAP', '0') != '1'
    plan = self.planner if cfg.transformer_decoder_join else self.planner_mlp
    if overlap:
      main, side = torch.cuda.current_stream(), self.side_stream(rgb.device)
      side.wait_stream(main)
      first = len(self.tape) if self.tape is not None else 0
      with torch.cuda.stream(side):
        pred_checkpoint, pred_target_speed, pred_wp = plan(fused, tp, ev_, cmd, training)
      if self.tape is not None:
        for r in self.tape[first:]:
          r['side'] = 'planner'
    else:
      pred_checkpoint, pred_target_speed, pred_wp = plan(fused, tp, ev_, cmd, training)
    pred_semantic = pred_depth = pred_bev_semantic = pred_bounding_box = None
    if cfg.use_semantic:
      pred_semantic = self.perspective_decoder(m.semantic_decoder, grid)
    if cfg.use_depth:
      pred_depth = self.perspective_decoder(m.depth_decoder, grid, ACT_SIGMOID).squeeze(1)
    if cfg.use_bev_semantic:
      dec = m.bev_semantic_decoder
      x = self.conv_bias(feats, dec[0], ACT_RELU)
      x = self.conv_bias(x, dec[2])
      ncls = dec[2].weight.shape[0]
      pred_bev_semantic = ops.bilinear_nchw_mask(x, ncls, cfg.lidar_resolution_height, cfg.lidar_resolution_width,
                                                 packed(m.valid_bev_pixels, 'f32'))
      self._save(op='bev_tail', src=x, out=pred_bev_semantic, ncls=ncls)
    if cfg.detect_boxes:
      pred_bounding_box = self.center_head_forward(feats)
    if overlap:
      main.wait_stream(side)
    return (pred_wp, pred_target_speed, pred_checkpoint, pred_semantic, pred_bev_semantic, pred_depth,
            pred_bounding_box, None, None, None)
