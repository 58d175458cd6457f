"""Training step of TransFuser++ on the libtfpp.so kernels: fused losses, hand-scheduled backward over the engine's
tape, flat fp32 master parameters with a fused AdamW(amsgrad) step and a bucketed NCCL all-reduce of the flat
gradient.  Replaces Engine.train's inner loop (team_code/train.py:883-916): forward, sum of weighted losses
(train.py:452-456,889-896), loss.backward() (train.py:898), optimizer.step() (train.py:908) and DDP's gradient
all-reduce (train.py:516).  Dropout (embd/attn/resid_pdrop = 0.1 and nn.TransformerDecoderLayer's 0.1) is not applied
— see DESIGN.md "Dropout".
"""
import os

import torch

from . import engine as eng_mod
from . import ops
from .engine import packed
from .ops import ACT_NONE, ACT_RELU, BF16, F32

LOSS_KEYS = ('loss_target_speed', 'loss_checkpoint', 'loss_semantic', 'loss_bev_semantic', 'loss_depth',
             'loss_center_heatmap', 'loss_wh', 'loss_offset', 'loss_yaw_class', 'loss_yaw_res')


def loss_keys(cfg):
  """Order of the loss VECTORS (Trainer graph output, autograd boundary): the ten default losses, then the optional
  ones.  The reference's dict (model.py:399-443) is keyed by name, so the order is ours to pick."""
  keys = LOSS_KEYS if cfg.use_controller_input_prediction else LOSS_KEYS[2:]
  return keys + (('loss_wp',) if getattr(cfg, 'use_wp_gru', False) else ())


def _pad8(n):
  return n + ((-n) % 8)


# ---------------------------------------------------------------------------------------------------------------
# flat parameter / gradient / optimizer state
# ---------------------------------------------------------------------------------------------------------------
class FlatState:
  """All trainable parameters re-pointed into ONE fp32 buffer (+ one gradient buffer, + AdamW state buffers).

  Adjacency groups make fused weight-gradient GEMMs write straight into the parameter gradients: (query, key, value)
  weights / biases of every fusion attention, and the five CenterNet head convs (3x3 weights, 3x3 biases, 1x1 biases).
  """

  def __init__(self, model, assign_grad=True, include_frozen=False, alloc=None):
    """alloc: optional callable total -> (flat, grad) zero-filled fp32 tensors of ``total`` elements (the NVLink peer
    buffers of carla_garage_b200.peer when the ranks exchange gradients through peer memory).  assign_grad: point every p.grad at its slice of the flat gradient (Trainer).  The autograd boundary leaves
    p.grad to torch's AccumulateGrad nodes instead.  include_frozen: also adopt requires_grad=False parameters
    (train.py:495-508 freezes sub-modules) so the backward handlers have somewhere to write; constants such as
    valid_bev_pixels stay out."""
    self.model = model
    constants = ('valid_bev_pixels', 'valid_bev_pixels_inv', 'grid', 'bev_projection_normalizer')  # geometry, not weights
    params = [(n, p) for n, p in model.named_parameters()
              if p.requires_grad or (include_frozen and p.is_floating_point() and n.rsplit('.', 1)[-1] not in constants)]
    by_name = dict(params)
    ordered, used = [], set()

    starts = set()  # first parameter of every adjacency group / every ungrouped parameter: may be preceded by padding

    def take(names, group=True):
      first = True
      for n in names:
        if n in by_name and n not in used:
          ordered.append((n, by_name[n]))
          used.add(n)
          if first or not group:
            starts.add(n)
          first = False

    for i in range(len(getattr(model.backbone, 'transformers', ()))):  # (the bev_encoder backbone has no fusion GPTs)
      for l in range(len(model.backbone.transformers[i].blocks)):
        base = f'backbone.transformers.{i}.blocks.{l}.attn.'
        take([base + 'query.weight', base + 'key.weight', base + 'value.weight'])
        take([base + 'query.bias', base + 'key.bias', base + 'value.bias'])
    if hasattr(model, 'head'):
      heads = model.head.head_names()
      take([f'head.{h}.0.weight' for h in heads])
      take([f'head.{h}.0.bias' for h in heads])
      take([f'head.{h}.2.bias' for h in heads])
    take([n for n, _ in params], group=False)
    self.names = [n for n, _ in ordered]
    self.params = [p for _, p in ordered]
    # matrices start on 16-byte boundaries (vector reductions / loads in the weight-gradient and pack kernels); the
    # padding floats stay zero forever (zero gradient, zero AdamW update)
    offs, off = [], 0
    for n, p in ordered:
      if n in starts and p.ndim >= 2:
        off += (-off) % 4
      offs.append(off)
      off += p.numel()
    total = off + ((-off) % 4)
    dev = self.params[0].device
    if alloc is not None:
      self.flat, self.grad = alloc(total)
      assert self.flat.numel() == total and self.grad.numel() == total and self.flat.dtype == F32
    else:
      self.flat = torch.zeros(total, dtype=F32, device=dev)
      self.grad = torch.zeros(total, dtype=F32, device=dev)
    self.offsets = {}
    for (n, p), off in zip(ordered, offs):
      k = p.numel()
      self.flat[off:off + k].copy_(p.detach().reshape(-1))
      p.data = self.flat[off:off + k].view(p.shape)
      if assign_grad:
        p.grad = self.grad[off:off + k].view(p.shape)
      self.offsets[id(p)] = (off, k)
    # every BatchNorm's num_batches_tracked as a view of one int64 buffer: one increment kernel per step
    counters = [(n, b) for n, b in model.named_buffers() if n.endswith('num_batches_tracked')]
    self.batch_counters = torch.zeros(max(len(counters), 1), dtype=torch.int64, device=dev)
    self.counter_index = {}
    for i, (_, b) in enumerate(counters):
      self.batch_counters[i] = b
      b.data = self.batch_counters[i]
      self.counter_index[id(b)] = i
    self.counter_mask = None  # 0/1 per counter: which BatchNorms a training forward actually runs (learned on step 1)
    self.exp_avg = torch.zeros_like(self.flat)
    self.exp_avg_sq = torch.zeros_like(self.flat)
    self.max_exp_avg_sq = torch.zeros_like(self.flat)
    self.step_count = 0
    # [step, lr, 1 - beta1^step, sqrt(1 - beta2^step)]: maintained on the device (CUDA-graph replay)
    self.dev_state = torch.zeros(4, dtype=F32, device=dev) if dev.type == 'cuda' else None
    self.flags = None  # per-element uint8 (bit 0 no weight decay, bit 1 frozen) or None = decay everything
    if any(not p.requires_grad for p in self.params):
      self.set_flags()

  def set_flags(self, no_decay=()):
    """Per-element optimizer flags for the fused AdamW: ``no_decay`` = parameters of the weight_decay=0 group of
    LidarCenterNet.create_optimizer_groups (model.py:556-645, train.py:522-525); requires_grad=False parameters
    (train.py:495-508) are frozen."""
    flags = torch.zeros(self.flat.numel(), dtype=torch.uint8, device=self.flat.device)
    nd = {id(p) for p in no_decay}
    for p in self.params:
      off, k = self.offsets[id(p)]
      v = (1 if id(p) in nd else 0) | (0 if p.requires_grad else 2)
      if v:
        flags[off:off + k] = v
    self.flags = flags if bool(flags.any()) else None

  def g(self, p):
    """fp32 gradient view of parameter p (same shape)."""
    off, k = self.offsets[id(p)]
    return self.grad[off:off + k].view(p.shape)

  def g_span(self, first, last):
    """contiguous gradient region from parameter ``first`` to parameter ``last`` (adjacency group)."""
    o0, _ = self.offsets[id(first)]
    o1, k1 = self.offsets[id(last)]
    return self.grad[o0:o1 + k1]

  def zero_grad(self):
    self.grad.zero_()

  def learn_counter_mask(self, bns):
    mask = torch.zeros_like(self.batch_counters)
    for bn in bns:
      mask[self.counter_index[id(bn.num_batches_tracked)]] += 1
    self.counter_mask = mask

  def count_batches(self):
    """one training forward: every BatchNorm it runs counts one more batch"""
    self.batch_counters += self.counter_mask

  def adamw_step(self, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, grad_scale=1.0):
    """optim.AdamW(amsgrad=True).step() (train.py:527-531,908) as one fused kernel over the flat buffers."""
    from . import _lib  # pylint: disable=import-outside-toplevel
    self.step_count += 1
    if lr is not None:  # None: keep the learning rate already on the device (CUDA-graph replay)
      self.dev_state[1:2].fill_(lr)
    _lib.check(_lib.load().tfpp_adamw_amsgrad(self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                              self.exp_avg_sq.data_ptr(), self.max_exp_avg_sq.data_ptr(),
                                              self.flat.numel(), 0.0, betas[0], betas[1], eps, weight_decay,
                                              self.step_count, grad_scale, self.dev_state.data_ptr(),
                                              ops._p(self.flags), ops._stream()),  # pylint: disable=protected-access
               'tfpp_adamw_amsgrad')
    eng_mod.PARAM_EPOCH[0] += 1  # parameter storage changed: cached bf16 weight packs must be rebuilt


# ---------------------------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------------------------
def loss_bias_targets(eng, st):
  """Where the fused loss kernels accumulate the bias gradients of the last conv of each dense head."""
  m = eng.m
  heads = m.head.head_names()
  return {'semantic': st.g(m.semantic_decoder.deconv3[2].bias), 'depth': st.g(m.depth_decoder.deconv3[2].bias),
          'center': st.g_span(getattr(m.head, heads[0])[2].bias, getattr(m.head, heads[-1])[2].bias)}


def _base_of(t):
  return t._base if t._base is not None else t  # pylint: disable=protected-access


def compute_losses(eng, outputs, labels, weights=None, bias_grads=None, w_dev=None, want_seeds=True):
  """Fused loss (+ seed-gradient) kernels: model.py:394-445 + center_net.py:77-123.

  Returns (dict of 10 loss scalars (0-dim device tensors), seeds).  ``weights``: dict loss key -> float multiplied into
  the seed gradients (train.py:452-456: 1/10 each; None = 1).  ``w_dev``: optional (10,) f32 device tensor in LOSS_KEYS
  order multiplied in as well, read by the kernels on the device (autograd boundary: d total / d loss_k).
  ``bias_grads``: {'semantic','depth','center'} -> fp32 tensors the kernels accumulate the last convs' bias gradients
  into.  want_seeds=False: loss values only.  seeds: key -> gradient of the (weighted) total wrt the pre-activation of a
  head, in the layout the backward GEMMs consume."""
  from . import _lib  # pylint: disable=import-outside-toplevel
  lib = _lib.load()
  m, cfg = eng.m, eng.cfg
  if any(float(w) != 1.0 for w in list(cfg.semantic_weights) + list(cfg.bev_semantic_weights)):
    raise NotImplementedError('tfpp_ce_map_loss assumes all-ones class weights on the semantic maps (config.py:163-164)')
  keys = loss_keys(cfg)
  weights = weights or {k: 1.0 for k in keys}
  bias_grads = bias_grads or {}
  pred_wp, pred_ts, pred_cp, pred_sem, pred_bev, pred_depth, bb = outputs[:7]
  dev = pred_sem.device
  b = pred_sem.shape[0]
  sums = eng.zeros((16,), dev)
  stream = ops._stream()  # pylint: disable=protected-access
  seeds = {}
  losses = {}
  # w_dev follows loss_keys(cfg); the kernels take pointers to the entry (or the first of two adjacent entries)
  wp = (lambda k: w_dev[keys.index(k):].data_ptr()) if w_dev is not None else (lambda k: None)
  if pred_wp is not None:  # waypoint L1 of the use_wp_gru branch (model.py:409-411): the L1 half of tfpp_planner_loss
    pred_wp = pred_wp.contiguous()
    dwp = torch.empty_like(pred_wp)
    wpt = wp('loss_wp')
    _lib.check(lib.tfpp_planner_loss(None, None, None, pred_wp.data_ptr(), labels['waypoint'].data_ptr(), 0.0,
                                     weights['loss_wp'], None if wpt is None else wpt - 4, sums[10:12].data_ptr(), None,
                                     dwp.data_ptr(), b, 0, pred_wp.shape[1] * pred_wp.shape[2], stream), 'wp_loss')
    losses['loss_wp'] = sums[11]
    seeds['planner_wp'] = (dwp, None)
  if pred_ts is not None:
    # target speed CE + checkpoint L1 (model.py:416-420)
    pred_ts, pred_cp = pred_ts.contiguous(), pred_cp.contiguous()
    dlogits = torch.empty_like(pred_ts)
    dcp = torch.empty_like(pred_cp)
    _lib.check(lib.tfpp_planner_loss(pred_ts.data_ptr(), labels['target_speed'].data_ptr(),
                                     packed(m.loss_speed.weight, 'f32').data_ptr(), pred_cp.data_ptr(),
                                     labels['checkpoint'].data_ptr(), weights['loss_target_speed'],
                                     weights['loss_checkpoint'], wp('loss_target_speed'), sums[0:2].data_ptr(),
                                     dlogits.data_ptr(), dcp.data_ptr(), b, pred_ts.shape[1],
                                     pred_cp.shape[1] * pred_cp.shape[2], stream), 'planner_loss')
    losses['loss_target_speed'], losses['loss_checkpoint'] = sums[0], sums[1]
    seeds['planner'] = (dcp, dlogits)
  # semantic CE (model.py:423)
  hw = pred_sem.shape[2] * pred_sem.shape[3]
  ncls = pred_sem.shape[1]
  cp = 16  # multiple of 16: the small-channel dgrad / wgrad kernels take 16- or 32-channel gradients
  dz = torch.empty((b, pred_sem.shape[2], pred_sem.shape[3], cp), dtype=ops.act_dtype(), device=dev) if want_seeds else None
  _lib.check(lib.tfpp_ce_map_loss(pred_sem.data_ptr(), labels['semantic'].data_ptr(), None,
                                  weights['loss_semantic'] / (b * hw), wp('loss_semantic'), sums[2:3].data_ptr(), ops._p(dz), None,  # pylint: disable=protected-access
                                  ops._p(bias_grads.get('semantic')), b, ncls, cp, hw, stream), 'ce semantic')  # pylint: disable=protected-access
  losses['loss_semantic'] = sums[2] / (b * hw)
  seeds['semantic'] = dz
  # BEV semantic CE with the frustum mask as ignore_index (model.py:426-431)
  hwb = pred_bev.shape[2] * pred_bev.shape[3]
  nb = pred_bev.shape[1]
  valid = packed(m.valid_bev_pixels, 'f32')
  n_valid = eng._const('n_valid_bev', lambda: m.valid_bev_pixels.detach().sum().cpu(), 'cpu')  # pylint: disable=protected-access
  count = float(n_valid) * b
  dbev = torch.empty_like(pred_bev) if want_seeds else None
  _lib.check(lib.tfpp_ce_map_loss(pred_bev.data_ptr(), labels['bev_semantic'].data_ptr(), valid.data_ptr(),
                                  weights['loss_bev_semantic'] / count, wp('loss_bev_semantic'), sums[3:4].data_ptr(), None, ops._p(dbev),  # pylint: disable=protected-access
                                  None, b, nb, 16, hwb, stream), 'ce bev')
  losses['loss_bev_semantic'] = sums[3] / count
  seeds['bev'] = dbev
  # depth L1 on the sigmoid output (model.py:379,434)
  n = pred_depth.numel()
  dzd = torch.empty((b, pred_depth.shape[-2], pred_depth.shape[-1], 16), dtype=ops.act_dtype(), device=dev) if want_seeds else None
  _lib.check(lib.tfpp_l1_sigmoid_loss(pred_depth.data_ptr(), labels['depth'].data_ptr(), weights['loss_depth'] / n,
                                      wp('loss_depth'), sums[4:5].data_ptr(), ops._p(dzd), ops._p(bias_grads.get('depth')), 16, n,  # pylint: disable=protected-access
                                      stream), 'l1 depth')
  losses['loss_depth'] = sums[4] / n
  seeds['depth'] = dzd
  # CenterNet head losses (center_net.py:77-123)
  maps = bb if torch.is_tensor(bb) else _base_of(bb[0])  # the fused (B,21,64,64) buffer
  hwc = maps.shape[2] * maps.shape[3]
  w5 = eng._const('w5_' + repr(sorted(weights.items())), lambda: torch.tensor(  # pylint: disable=protected-access
      [weights[k] for k in ('loss_center_heatmap', 'loss_wh', 'loss_offset', 'loss_yaw_class', 'loss_yaw_res')],
      dtype=F32), dev)
  if w_dev is not None:
    i5 = keys.index('loss_center_heatmap')
    w5 = w5 * w_dev[i5:i5 + 5]
  dzh = torch.empty((b, maps.shape[2], maps.shape[3], 24), dtype=ops.act_dtype(), device=dev) if want_seeds else None
  _lib.check(lib.tfpp_center_head_loss(maps.data_ptr(), labels['center_heatmap'].data_ptr(), labels['wh'].data_ptr(),
                                       labels['offset'].data_ptr(), labels['yaw_class'].data_ptr(),
                                       labels['yaw_res'].data_ptr(), labels['pixel_weight'].data_ptr(),
                                       labels['avg_factor'].data_ptr(), w5.data_ptr(), sums[5:10].data_ptr(),
                                       ops._p(dzh), ops._p(bias_grads.get('center')), b, hwc,  # pylint: disable=protected-access
                                       cfg.num_bb_classes, cfg.num_dir_bins, 24, stream), 'center loss')
  for i, k in enumerate(('loss_center_heatmap', 'loss_wh', 'loss_offset', 'loss_yaw_class', 'loss_yaw_res')):
    losses[k] = sums[5 + i]
  seeds['center'] = dzh
  return losses, seeds


# ---------------------------------------------------------------------------------------------------------------
# backward over the tape
# ---------------------------------------------------------------------------------------------------------------
class Backward:

  def __init__(self, eng, st):
    self.eng, self.st = eng, st
    self.G = {}        # id(forward tensor) -> gradient tensor (bf16 NHWC / f32 token matrices)
    self.pending = {}  # id(a2) -> (gate, pool_grad) from the SE record

  def add(self, t, g):
    k = id(t)
    if k in self.G:
      if g.dtype == BF16:
        ops.add_bf16(self.G[k], g, out=self.G[k])
      else:
        self.G[k].add_(g)
    else:
      self.G[k] = g

  # -- generic pieces ------------------------------------------------------------------------------------------
  def linear_bwd(self, dy_b, x, w_param_grad, wt_packed, n, k, *, res=None, out_f32=True, need_dx=True, cout_valid=0,
                 x_slab=None):
    """dy_b (rows, Np) bf16, x (rows, K) bf16 [or slab], weight grad region (n, k) fp32 contiguous.
    Returns dx (rows, k) (f32 by default, + res)."""
    rows = dy_b.shape[0]
    npad = dy_b.shape[1]
    if x_slab is None:
      ops.conv_wgrad(dy_b.view(1, 1, rows, npad), x.view(1, 1, rows, k), out=w_param_grad, out_strides=(k, 0, 1),
                     cout_valid=cout_valid or n)
    else:
      buf, off, (groups, rows_pg), gstride = x_slab
      ops.conv_wgrad(dy_b.view(groups, 1, rows_pg, npad), buf.view(-1)[off:], x_shape=(groups, 1, rows_pg, k),
                     x_batch_stride=gstride, out=w_param_grad, out_strides=(k, 0, 1), cout_valid=cout_valid or n)
    if not need_dx:
      return None
    return ops.linear(dy_b, wt_packed, res=res, out_f32=out_f32)

  def conv_dgrad(self, dz, conv_weight, taps_kind, shape_out, res=None):
    """dz (B,H,W,Cp) bf16 -> grad wrt the conv input (B,H,W,Cin) bf16 for a dense stride-1 conv."""
    wt = packed(conv_weight, 'conv_t')
    taps = ops.TAPS_3X3_DGRAD if taps_kind == 3 else ops.TAPS_1X1
    return ops.conv_gemm(dz, wt, taps=taps, res1=res)

  # -- handlers ------------------------------------------------------------------------------------------------
  def conv_bias(self, r, seeds):
    st = self.st
    conv, y, a, act = r['conv'], r['y'], r['a'], r['act']
    cout, cin = conv.weight.shape[0], conv.weight.shape[1]
    k = conv.weight.shape[-1]
    cp = _pad8(cout)
    b, h, w = a.shape[0], a.shape[1], a.shape[2]
    if id(y) in seeds:  # final output: the loss kernel produced dz (and the bias gradient)
      dz = seeds.pop(id(y))
    else:
      dy = self.G.pop(id(y))
      if r['kw'].get('out_layout') == 'nchw':  # NCHW f32 head output consumed by something other than a loss
        raise RuntimeError('unexpected NCHW f32 intermediate')
      if r.get('smallc'):
        cp = 16 if cout <= 16 else 32
      if act == ACT_NONE and cp == cout:
        dz = dy
        ops.act_bwd(dy, None, ACT_NONE, b, h * w, cout, dbias=st.g(conv.bias), want_dz=False)
      else:
        dz = ops.act_bwd(dy, y, act, b, h * w, cout, dbias=st.g(conv.bias), channels_padded=cp).view(b, h, w, cp)
    dz = dz.view(b, h, w, -1)
    cp = dz.shape[3]
    if r.get('smallc') and cp in (16, 32):
      ops.smallc_wgrad3x3(dz, a, st.g(conv.weight), (cin * 9, 1, 9), cout)
      if r.get('halo') and ops.halo_umma_supported(cp, cin):  # experimental tcgen05 path (TFPP_HALO_UMMA=1)
        da = ops.halo_conv3x3(dz, packed(conv.weight, 'conv_halo_umma_t', cp))
      else:
        da = ops.smallc_conv3x3(dz, packed(conv.weight, 'conv_dgrad_smallc', cp))
      if id(a) in self.G:
        ops.add_bf16(self.G[id(a)], da, out=da)
      self.G[id(a)] = da
      return
    ops.conv_wgrad(dz, a, cin=cin, taps=r['taps'], w_taps=k * k, out=st.g(conv.weight),
                   out_strides=(cin * k * k, 1, k * k), cout_valid=cout)
    self.G[id(a)] = self.conv_dgrad(dz, conv.weight, k, None, res=self.G.get(id(a)))

  def conv_bn(self, r):
    st = self.st
    cna, a, raw, y, act = r['cna'], r['a'], r['raw'], r['y'], r['act']
    bn, conv = cna.bn, cna.conv
    dy = self.G.pop(id(y))
    gate = pool_grad = None
    if id(y) in self.pending:
      gate, pool_grad = self.pending.pop(id(y))
    want_dz = r['res'] is not None or r['res_bn'] is not None
    # without a residual the ReLU mask is a function of raw alone: recompute it instead of reading y
    from_raw = act == ACT_RELU and not want_dz
    draw, dz = ops.bn_bwd(dy, None if from_raw else y, raw, r['mean'], r['invstd'], bn.weight, act, st.g(bn.weight),
                          st.g(bn.bias), gate=gate, pool_grad=pool_grad, want_dz=want_dz,
                          fwd_affine=(r['scale'], r['shift']) if from_raw else None)
    if r['res'] is not None:
      self.add(r['res'], dz)
    if r['res_bn'] is not None:
      self.G[id(r['res_bn'][0])] = dz
    self.conv_grads(draw, a, r['a_src'], conv, r['taps'], r['batch'], r['grouped'], r.get('stride', 1))

  def conv_grads(self, draw, a, a_src, conv, taps, batch, grouped, stride=1):
    """weight gradient + input gradient of a bias-free conv given the gradient of its raw output."""
    st = self.st
    k = conv.weight.shape[-1]
    if grouped:  # RegNet 3x3 group conv; a = the full-resolution input whatever the stride
      gout = st.g(conv.weight)
      ops.gconv3x3_wgrad(draw, a, gout, stride)
      wt = packed(conv.weight, 'gconv_halo_t')  # transposed + spatially flipped weights
      if stride == 1 and eng_mod.HALO_UMMA_GCONV:  # experimental tcgen05 path (TFPP_HALO_UMMA_GCONV=1)
        da = ops.halo_gconv3x3(draw, packed(conv.weight, 'gconv_halo_umma_t'))
      else:
        da = ops.gconv3x3(draw, wt) if stride == 1 else ops.gconv3x3_dgrad_s2(draw, wt)
      if id(a) in self.G:
        ops.add_bf16(self.G[id(a)], da, out=da)
      self.G[id(a)] = da
    else:
      cin = conv.weight.shape[1]
      ops.conv_wgrad(draw, a, cin=cin, taps=taps, w_taps=k * k, out=st.g(conv.weight),
                     out_strides=(cin * k * k, 1, k * k))
      wt = packed(conv.weight, 'conv_t')
      if a_src is None:
        tp = ops.TAPS_3X3_DGRAD if k == 3 else ops.TAPS_1X1
        self.G[id(a)] = ops.conv_gemm(draw, wt, taps=tp, res1=self.G.get(id(a)))
      else:  # 1x1 stride 2 (downsample): only the even/even positions receive gradient
        h, w, c = a_src.shape[1], a_src.shape[2], a_src.shape[3]
        da = self.G.get(id(a_src))
        have = da is not None
        if not have:
          da = torch.zeros_like(a_src)
        strides = (h * w * c, 2 * w * c, 2 * c, 1)
        ops.conv_gemm(draw, wt, out=da, out_strides=strides, res1=da, res1_strides=strides)
        self.G[id(a_src)] = da

  def downsample(self, r):
    st = self.st
    cna = r['cna']
    dz = self.G.pop(id(r['raw']))
    draw, _ = ops.bn_bwd(dz, None, r['raw'], r['mean'], r['invstd'], cna.bn.weight, ACT_NONE, st.g(cna.bn.weight),
                         st.g(cna.bn.bias))
    self.conv_grads(draw, r['a'], r['x_src'] if r['stride'] == 2 else None, cna.conv, ops.TAPS_1X1, r['batch'], False)

  def se(self, r):
    st = self.st
    se = r['se']
    da2s = self.G.pop(id(r['a2s']))
    pool_grad = ops.se_bwd(da2s, r['a2'], r['gate'], r['hidden'], r['pool'], r['hw'], se.fc1.weight, se.fc2.weight,
                           st.g(se.fc1.weight), st.g(se.fc1.bias), st.g(se.fc2.weight), st.g(se.fc2.bias),
                           zeros=self.eng.zeros)
    self.pending[id(r['a2'])] = (r['gate'], pool_grad)
    self.G[id(r['a2'])] = da2s

  def stem(self, r):
    st = self.st
    cna = r['cna']
    dy = self.G.pop(id(r['y']), None)
    if dy is None:
      return
    draw, _ = ops.bn_bwd(dy, None, r['raw'], r['mean'], r['invstd'], cna.bn.weight, ACT_RELU, st.g(cna.bn.weight),
                         st.g(cna.bn.bias), fwd_affine=(r['scale'], r['shift']))
    ops.stem_wgrad(r['x'], draw, r['in_scale'], r['in_shift'], st.g(cna.conv.weight))

  def bilinear(self, r):
    src, out = r['src'], r['out']
    dout = self.G.pop(id(out))
    b, sh, sw, c = src.shape
    dsrc = self.G.get(id(src))
    have = dsrc is not None
    if not have:
      dsrc = torch.empty_like(src)
    ops.bilinear_bwd(dout, dsrc, b, sh, sw, out.shape[1], out.shape[2], c, accumulate=have)
    self.G[id(src)] = dsrc

  def untokenise(self, r):
    """adjoint of: img_out = img + up(xf[:, :n_img]); lid_out = lid + up(conv1x1(xf[:, n_img:]))."""
    st, eng = self.st, self.eng
    bb, cfg = eng.bb, eng.cfg
    i, b, t, c, cl = r['i'], r['b'], r['t'], r['c'], r['cl']
    gpt = bb.transformers[i]
    n_img = cfg.img_vert_anchors * cfg.img_horz_anchors
    n_lid = t - n_img
    dimg = self.G.pop(id(r['img_out']))
    dlid = self.G.pop(id(r['lid_out']))
    # pass-through of the residual adds
    self.add(r['img'], dimg)
    self.add(r['lid'], dlid)
    dxf = torch.empty((b * t, c), dtype=F32, device=dimg.device)
    hi, wi = dimg.shape[1], dimg.shape[2]
    ops.bilinear_bwd(dimg, dxf, b, cfg.img_vert_anchors, cfg.img_horz_anchors, hi, wi, c, src_batch_stride=t * c,
                     src_row_stride=c)
    dlid_tok = torch.empty((b, cfg.lidar_vert_anchors, cfg.lidar_horz_anchors, cl), dtype=ops.act_dtype(), device=dimg.device)
    ops.bilinear_bwd(dlid, dlid_tok, b, cfg.lidar_vert_anchors, cfg.lidar_horz_anchors, dlid.shape[1], dlid.shape[2], cl)
    i2l = bb.img_channel_to_lidar[i]
    dlt = dlid_tok.view(b * n_lid, cl)
    ops.act_bwd(dlt, None, ACT_NONE, 1, b * n_lid, cl, dbias=st.g(i2l.bias), want_dz=False)
    # weight gradient: x = xf slab rows n_img.. of every sample
    xf = r['xf']
    ops.conv_wgrad(dlt.view(b, 1, n_lid, cl), xf.view(-1)[n_img * c:], x_shape=(b, 1, n_lid, c), x_batch_stride=t * c,
                   out=st.g(i2l.weight), out_strides=(c, 0, 1))
    # dxf[:, n_img:] = dlt @ W_i2l  (written into the slab)
    ops.linear(dlt, packed(i2l.weight, 'linear_t'), out=dxf.view(-1)[n_img * c:], row_map=(n_lid, t))
    dx = ops.layernorm_bwd(dxf, r['x'], r['meanf'], r['rstdf'], gpt.ln_f.weight, st.g(gpt.ln_f.weight),
                           st.g(gpt.ln_f.bias))
    self.G[id(r['x'])] = dx

  def gpt_block(self, r):
    st = self.st
    blk, b, t, c = r['blk'], r['b'], r['t'], r['c']
    at = blk.attn
    rows = b * t
    dx2 = self.G.pop(id(r['x2']))
    # x2 = x1 + mlp2(m); m = relu(mlp1(h2)); h2 = LN2(x1)
    l1, l2 = blk.mlp[0], blk.mlp[2]
    d_attn, d_proj, d_mlp = r.get('drops', (None, None, None))
    dz2 = ops.act_bwd(dx2, None, ACT_NONE, 1, rows, c, layout=2, dbias=st.g(l2.bias), drop=d_mlp)
    dm = self.linear_bwd(dz2, r['m'], st.g(l2.weight), packed(l2.weight, 'linear_t'), c, 4 * c, out_f32=False)
    dzm = ops.act_bwd(dm, r['m'], ACT_RELU, 1, rows, 4 * c, dbias=st.g(l1.bias))
    dh2 = self.linear_bwd(dzm, r['h2'], st.g(l1.weight), packed(l1.weight, 'linear_t'), 4 * c, c)
    dx1 = ops.layernorm_bwd(dh2, r['x1'], r['mean2'], r['rstd2'], blk.ln2.weight, st.g(blk.ln2.weight),
                            st.g(blk.ln2.bias), dres=dx2)
    # x1 = x + proj(y); y = attn(qkv); qkv = lin(h); h = LN1(x)
    dzp = ops.act_bwd(dx1, None, ACT_NONE, 1, rows, c, layout=2, dbias=st.g(at.proj.bias), drop=d_proj)
    dy = self.linear_bwd(dzp, r['y'], st.g(at.proj.weight), packed(at.proj.weight, 'linear_t'), c, c, out_f32=False)
    dqkv = ops.fusion_attn_bwd(r['qkv'], dy, b, t, c, r['heads'], drop=d_attn)
    ops.act_bwd(dqkv, None, ACT_NONE, 1, rows, 3 * c, dbias=st.g_span(at.query.bias, at.value.bias), want_dz=False)
    wt = packed((at.query.weight, at.key.weight, at.value.weight), 'cat_linear_t')
    dh = self.linear_bwd(dqkv, r['h'], st.g_span(at.query.weight, at.value.weight), wt, 3 * c, c)
    dx = ops.layernorm_bwd(dh, r['x'], r['mean1'], r['rstd1'], blk.ln1.weight, st.g(blk.ln1.weight),
                           st.g(blk.ln1.bias), dres=dx1)
    self.G[id(r['x'])] = dx

  def tokenise(self, r):
    st, eng = self.st, self.eng
    bb, cfg = eng.bb, eng.cfg
    i, b, t, c, cl = r['i'], r['b'], r['t'], r['c'], r['cl']
    gpt = bb.transformers[i]
    n_img = cfg.img_vert_anchors * cfg.img_horz_anchors
    n_lid = t - n_img
    dx = self.G.pop(id(r['x0']))
    ops.dropout_(dx, r.get('drop'))  # adjoint of GPT.drop (in place: dx is ours)
    # pos_emb gradient = sum over the batch
    ops.batch_reduce(dx, st.g(gpt.pos_emb).view(-1), b)
    img, lid = r['img'], r['lid']
    self.G[id(img)] = ops.pool_bwd_add(self.G.get(id(img)), dx, tuple(img.shape), cfg.img_vert_anchors,
                                       cfg.img_horz_anchors, t, 0)
    l2i = bb.lidar_channel_to_img[i]
    dzl = ops.cast_rows(dx, b, t, n_img, n_lid, c, dbias=st.g(l2i.bias))
    dpool = self.linear_bwd(dzl, r['lid_pool'].view(b * n_lid, cl), st.g(l2i.weight), packed(l2i.weight, 'linear_t'), c,
                            cl, out_f32=False)
    self.G[id(lid)] = ops.pool_bwd_add(self.G.get(id(lid)), dpool, tuple(lid.shape), cfg.lidar_vert_anchors,
                                       cfg.lidar_horz_anchors, n_lid, 0)

  def center_head(self, r, seeds):
    st = self.st
    feat, h = r['feat'], r['h']
    convs0, convs1 = r['convs0'], r['convs1']
    dz = seeds.pop('center')  # (B,64,64,24) bf16, 1x1-conv bias gradients already accumulated by the loss kernel
    b, hh, ww, _ = dz.shape
    n1 = sum(c.weight.shape[0] for c in convs1)
    k1 = sum(c.weight.shape[1] for c in convs1)
    tmp = ops.conv_wgrad(dz, h, cin=k1, cout_valid=n1)  # (21, 1, 320) fp32
    rr = cc = 0
    for c in convs1:
      n, k = c.weight.shape[0], c.weight.shape[1]
      st.g(c.weight).view(n, k).add_(tmp[rr:rr + n, 0, cc:cc + k])
      rr += n
      cc += k
    dh = ops.conv_gemm(dz, packed(tuple(c.weight for c in convs1), 'blockdiag_1x1_t'))
    n0 = k1
    dzh = ops.act_bwd(dh, h, ACT_RELU, b, hh * ww, n0, dbias=st.g_span(convs0[0].bias, convs0[-1].bias)).view(b, hh, ww, n0)
    cin = convs0[0].weight.shape[1]
    ops.conv_wgrad(dzh, feat, cin=cin, taps=ops.TAPS_3X3, w_taps=9, out=st.g_span(convs0[0].weight, convs0[-1].weight),
                   out_strides=(cin * 9, 1, 9))
    wt = packed(tuple(c.weight for c in convs0), 'cat_conv_t')
    self.G[id(feat)] = ops.conv_gemm(dzh, wt, taps=ops.TAPS_3X3_DGRAD, res1=self.G.get(id(feat)))

  def bev_tail(self, r, seeds):
    src, out = r['src'], r['out']
    dbev = seeds.pop('bev')
    b, sh, sw, _ = src.shape
    ncls = r['ncls']
    cp = _pad8(ncls)
    dz = ops.bilinear_nchw_mask_bwd(dbev, packed(self.eng.m.valid_bev_pixels, 'f32'), b, sh, sw, cp, ncls, out.shape[2],
                                    out.shape[3])
    # hand it to the 1x1 conv record as a seed; its bias gradient is the column sum of dz
    conv = self.eng.m.bev_semantic_decoder[2]
    tmpb = torch.zeros(cp, dtype=F32, device=dz.device)
    ops.act_bwd(dz, None, ACT_NONE, b, sh * sw, cp, dbias=tmpb, want_dz=False)
    self.st.g(conv.bias).add_(tmpb[:ncls])
    seeds[id(src)] = dz

  def planner_head(self, r, seeds):
    from . import _lib  # pylint: disable=import-outside-toplevel
    st, m = self.st, self.eng.m
    dcp, dlogits = seeds.pop(r.get('seed_key', 'planner'))
    b, nq, d = r['b'], r['nq'], r['d']
    cd, tsn = r['cd'], r['tsn']
    joined = r['joined']
    h_all = r['res'][2]
    djoined = torch.empty_like(joined)
    g = st.g
    tp = r['target_point'].float().contiguous()
    n_speed = tsn[2].weight.shape[0] if tsn is not None else 0
    n_wp = nq - 1 if n_speed else nq
    P = lambda t: None if t is None else t.data_ptr()
    tw = (tsn[0].weight, tsn[0].bias, tsn[2].weight) if tsn is not None else (None, None, None)
    tg = (g(tsn[0].weight), g(tsn[0].bias), g(tsn[2].weight), g(tsn[2].bias)) if tsn is not None else (None,) * 4
    _lib.check(_lib.load().tfpp_planner_head_bwd(
        joined.data_ptr(), tp.data_ptr(), h_all.data_ptr(), cd.encoder.weight.data_ptr(), cd.encoder.bias.data_ptr(),
        cd.gru.weight_ih_l0.data_ptr(), cd.gru.weight_hh_l0.data_ptr(), cd.gru.bias_ih_l0.data_ptr(),
        cd.gru.bias_hh_l0.data_ptr(), cd.decoder.weight.data_ptr(), P(tw[0]), P(tw[1]), P(tw[2]), dcp.data_ptr(),
        P(dlogits), djoined.data_ptr(), g(cd.encoder.weight).data_ptr(),
        g(cd.encoder.bias).data_ptr(), g(cd.gru.weight_ih_l0).data_ptr(), g(cd.gru.weight_hh_l0).data_ptr(),
        g(cd.gru.bias_ih_l0).data_ptr(), g(cd.gru.bias_hh_l0).data_ptr(), g(cd.decoder.weight).data_ptr(),
        g(cd.decoder.bias).data_ptr(), P(tg[0]), P(tg[1]), P(tg[2]), P(tg[3]), b, n_wp, d, cd.hidden_size, n_speed,
        ops._stream()), 'planner_head_bwd')  # pylint: disable=protected-access
    mj, rj = r['stats']
    dx = ops.layernorm_bwd(djoined.view(b * nq, d), r['x'], mj, rj, m.join.norm.weight, g(m.join.norm.weight),
                           g(m.join.norm.bias))
    self.G[id(r['x'])] = dx

  def gru_cell_head(self, r, seeds):
    """adjoint of Engine._gru_cell_head: BPTT through the autoregressive GRUCell (+ target-speed MLP)."""
    st, g = self.st, self.st.g
    dwp, dts = seeds.pop(r['seed_key'])
    cd, tsn, joined = r['cd'], r['tsn'], r['joined']
    cell = cd.wp_decoder
    dj = self.G.get(id(joined))
    if dj is None:   # the heads share the joined feature: the kernels accumulate into one zero-initialised gradient
      dj = torch.zeros_like(joined)
      self.G[id(joined)] = dj
    tw = (tsn[0].weight, tsn[0].bias, tsn[2].weight) if tsn is not None else (None, None, None)
    tg = (g(tsn[0].weight), g(tsn[0].bias), g(tsn[2].weight), g(tsn[2].bias)) if tsn is not None else (None,) * 4
    ops.gru_cell_head_bwd(joined, r['target_point'], cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh,
                          cd.output.weight, cd.output.bias, *tw, r['wp'], r['h_all'], dwp.contiguous(),
                          dts.contiguous() if (dts is not None and tsn is not None) else None, dj,
                          (g(cell.weight_ih), g(cell.weight_hh), g(cell.bias_ih), g(cell.bias_hh), g(cd.output.weight),
                           g(cd.output.bias)) + tg, steps=cd.prediction_len, hidden=cd.hidden_size,
                          learn_origin=bool(self.eng.cfg.learn_origin))

  def mlp_join(self, r):
    """adjoint of Engine.planner_mlp up to the joined feature (model.py:306-322,360)."""
    from . import _lib  # pylint: disable=import-outside-toplevel
    st, m, g = self.st, self.eng.m, self.st.g
    b, c, e = r['b'], r['c'], r['e']
    fused, es, h1, h2, joined = r['fused'], r['es'], r['h1'], r['h2'], r['joined']
    j0, j1, j2 = m.join[0], m.join[2], m.join[4]
    dj = self.G.pop(id(joined))
    n2 = j2.weight.shape[0]
    # joined is f32 (b, n2): with one pixel per sample the NCHW-f32 layout of act_bwd is exactly that matrix
    dz3 = ops.act_bwd(dj, joined, ACT_RELU, b, 1, n2, layout=1, dbias=g(j2.bias), channels_padded=_pad8(n2))
    dh2 = self.linear_bwd(dz3, h2, g(j2.weight), packed(j2.weight, 'linear_t'), n2, j2.weight.shape[1], out_f32=False,
                          cout_valid=n2)
    dz2 = ops.act_bwd(dh2, h2, ACT_RELU, 1, b, j1.weight.shape[0], dbias=g(j1.bias))
    dh1 = self.linear_bwd(dz2, h1, g(j1.weight), packed(j1.weight, 'linear_t'), j1.weight.shape[0], j1.weight.shape[1],
                          out_f32=False)
    n0 = j0.weight.shape[0]
    dz1 = ops.act_bwd(dh1, h1, ACT_RELU, 1, b, n0, dbias=g(j0.bias))
    gw0 = g(j0.weight)   # (256, c + e): the two column blocks receive their own weight gradients
    ops.conv_wgrad(dz1.view(1, 1, b, n0), fused.view(1, 1, b, c), out=gw0, out_strides=(c + e, 0, 1))
    ops.conv_wgrad(dz1.view(1, 1, b, n0), es.view(1, 1, b, e), out=gw0[:, c:], out_strides=(c + e, 0, 1))
    self.add(fused, ops.linear(dz1, packed(j0.weight, 'cols_t', 0, c)))
    des = ops.linear(dz1, packed(j0.weight, 'cols_t', c, c + e), out_f32=True)
    ese, vn = m.extra_sensor_encoder, m.velocity_normalization
    training = r['training']
    dpos = torch.zeros(e, dtype=F32, device=des.device)   # the kernel's positional-embedding slot: unused on this path
    _lib.check(_lib.load().tfpp_extra_sensor_token_bwd(
        r['ego_vel'].float().contiguous().data_ptr(), r['command'].float().contiguous().data_ptr(),
        0.0 if training else float(vn.running_mean[0]), 1.0 if training else float(vn.running_var[0]), int(training),
        ese[0].weight.data_ptr(), ese[0].bias.data_ptr(), ese[2].weight.data_ptr(), ese[2].bias.data_ptr(), des.data_ptr(), e,
        g(ese[0].weight).data_ptr(), g(ese[0].bias).data_ptr(), g(ese[2].weight).data_ptr(), g(ese[2].bias).data_ptr(),
        dpos.data_ptr(), b, r['command'].shape[1], ese[0].weight.shape[0], e, ops._stream()), 'extra_sensor_bwd')  # pylint: disable=protected-access

  def global_fuse(self, r):
    """adjoint of the global pools + lidar_to_img_features_end + sum (transfuser.py:188-197)."""
    st, bb, g = self.st, self.eng.bb, self.st.g
    b, c, cl = r['b'], r['c'], r['cl']
    img, lid = r['img'], r['lid']
    df = self.G.pop(id(r['fused']))   # (b, c)
    end = bb.lidar_to_img_features_end
    ops.act_bwd(df, None, ACT_NONE, 1, b, c, dbias=g(end.bias), want_dz=False)
    dlp = self.linear_bwd(df, r['lid_pool'].view(b, cl), g(end.weight), packed(end.weight, 'linear_t'), c, cl, out_f32=False)
    self.G[id(img)] = ops.pool_bwd_add(self.G.get(id(img)), df, tuple(img.shape), 1, 1, 1, 0)
    self.G[id(lid)] = ops.pool_bwd_add(self.G.get(id(lid)), dlp, tuple(lid.shape), 1, 1, 1, 0)

  def conv_in(self, r):
    """adjoint of Engine.conv_in: InstanceNorm2d + activation, then weight / input gradients of every Cin slice."""
    st = self.st
    conv, raw = r['conv'], r['raw']
    dy = self.G.pop(id(r['y']))
    draw = ops.instnorm_bwd(dy, raw, r['mean'], r['invstd'], r['act'], dy_pix_stride=r['y_pix_stride'], zeros=self.eng.zeros)
    cin_all, k = conv.weight.shape[1], conv.weight.shape[-1]
    gw = st.g(conv.weight).view(-1)
    for a, k0, k1 in r['slices']:
      ops.conv_wgrad(draw, a, cin=k1 - k0, taps=ops.TAPS_3X3, w_taps=k * k, out=gw[k0 * k * k:],
                     out_strides=(cin_all * k * k, 1, k * k))
      self.G[id(a)] = ops.conv_gemm(draw, packed(conv.weight, 'conv_cin_t', k0, k1), taps=ops.TAPS_3X3_DGRAD,
                                    res1=self.G.get(id(a)))

  def bev_lift(self, r):
    img = r['img']
    dout = self.G.pop(id(r['out']))
    self.G[id(img)] = ops.bev_lift_bwd(dout, r['tables'], tuple(img.shape), dimg=self.G.get(id(img)))

  def bev_stem(self, r):
    """adjoint of Engine.bev_stem: BatchNorm + ReLU, the weight gradient over the parity planes, and the input gradient
    as one implicit GEMM per parity plane written straight to its pixels of the full-resolution tensor."""
    st = self.st
    cna, cat, planes = r['cna'], r['cat'], r['planes']
    conv, bn = cna.conv, cna.bn
    dy = self.G.pop(id(r['y']))
    draw, _ = ops.bn_bwd(dy, None, r['raw'], r['mean'], r['invstd'], bn.weight, ACT_RELU, st.g(bn.weight), st.g(bn.bias),
                         fwd_affine=(r['scale'], r['shift']))
    b, h, w, cpad = cat.shape
    cout, cin = conv.weight.shape[0], conv.weight.shape[1]
    dw = ops.conv_wgrad(draw, planes, cin=cpad, taps=ops.taps_3x3_stride2(b), w_taps=9)      # (Cout, 9, cpad) f32
    st.g(conv.weight).view(cout, cin, 9).add_(dw[:, :, :cin].transpose(1, 2))
    wt = packed(conv.weight, 'conv_cin_pad_t', cpad)                                          # (cpad, 9, Cout)
    dcat = torch.empty_like(cat)
    flat = dcat.view(-1)
    strides = (h * w * cpad, 2 * w * cpad, 2 * cpad, 1)
    for py in (0, 1):
      for px in (0, 1):
        ops.conv_gemm(draw, wt, taps=ops.taps_3x3_stride2_dgrad(py, px), out=flat[(py * w + px) * cpad:],
                      out_strides=strides)
    self.G[id(cat)] = dcat

  def planner_queries(self, r):
    # a learned query set repeated over the batch (model.py:329,349): its gradient is the sum over the batch
    dx0 = self.G.pop(id(r['x0']))
    ops.batch_reduce(dx0, self.st.g(r['query']).view(-1), r['b'])

  def dec_layer(self, r):
    st = self.st
    l, b, nq, n_mem, d, heads, hd = r['layer'], r['b'], r['nq'], r['n_mem'], r['d'], r['heads'], r['hd']
    g = st.g
    rows = b * nq
    m1, r1, m2, r2, m3, r3 = r['stats']
    dx3 = self.G.pop(id(r['x3']))
    # x3 = LN3(t3), t3 = x2 + lin2(ff), ff = act(lin1(x2b))
    dt3 = ops.layernorm_bwd(dx3, r['t3'], m3, r3, l.norm3.weight, g(l.norm3.weight), g(l.norm3.bias))
    dp = r.get('drops', (None,) * 6)
    dz = ops.act_bwd(dt3, None, ACT_NONE, 1, rows, d, layout=2, dbias=g(l.linear2.bias), drop=dp[5])
    ffw = l.linear1.weight.shape[0]
    dff = self.linear_bwd(dz, r['ff'], g(l.linear2.weight), packed(l.linear2.weight, 'linear_t'), d, ffw, out_f32=False)
    if r['act'] != ACT_RELU:
      raise NotImplementedError('GELU decoder feed-forward backward is not built (the reference runs ReLU)')
    # ff is stored AFTER its dropout: ff > 0 <=> kept and ReLU-active, so the stored tensor is the whole mask and the
    # dropout adjoint reduces to the 1/(1-p) factor
    dzf = ops.act_bwd(dff, r['ff'], ACT_RELU, 1, rows, ffw, dbias=g(l.linear1.bias),
                      dy_scale=1.0 / (1.0 - dp[4][1]) if dp[4] is not None else 1.0)
    dx2 = self.linear_bwd(dzf, r['x2b'], g(l.linear1.weight), packed(l.linear1.weight, 'linear_t'), ffw, d, res=dt3)
    # x2 = LN2(t2), t2 = x1 + out_proj(ca), ca = mha(q2(x1b), kv)
    dt2 = ops.layernorm_bwd(dx2, r['t2'], m2, r2, l.norm2.weight, g(l.norm2.weight), g(l.norm2.bias))
    mh = l.multihead_attn
    dz = ops.act_bwd(dt2, None, ACT_NONE, 1, rows, d, layout=2, dbias=g(mh.out_proj.bias), drop=dp[3])
    dca = self.linear_bwd(dz, r['ca'], g(mh.out_proj.weight), packed(mh.out_proj.weight, 'linear_t'), d, d,
                          out_f32=False)
    kv = r['kv']
    dq2 = torch.empty((rows, d), dtype=ops.act_dtype(), device=kv.device)
    dkv = torch.empty_like(kv)
    ops.small_mha_bwd(r['q2'], kv, kv, dca, dq2, dkv, dkv, b, heads, nq, n_mem, hd, (nq * d, d), (n_mem * 2 * d, 2 * d),
                      (n_mem * 2 * d, 2 * d), (nq * d, d), (n_mem * 2 * d, 2 * d), (n_mem * 2 * d, 2 * d),
                      offs=(0, 0, d, 0, 0, d), drop=dp[2])
    # q projection (rows [0,d) of in_proj) and k/v projections (rows [d,3d)) of the cross attention
    ipw, ipb = g(mh.in_proj_weight), g(mh.in_proj_bias)
    ops.act_bwd(dq2, None, ACT_NONE, 1, rows, d, dbias=ipb[:d], want_dz=False)
    dx1 = self.linear_bwd(dq2, r['x1b'], ipw[:d], packed(mh.in_proj_weight, 'rows_t', 0, d), d, d, res=dt2)
    ops.act_bwd(dkv, None, ACT_NONE, 1, b * n_mem, 2 * d, dbias=ipb[d:], want_dz=False)
    mem = self.eng_mem
    dmem = self.linear_bwd(dkv, mem, ipw[d:], packed(mh.in_proj_weight, 'rows_t', d, 3 * d), 2 * d, d,
                           res=self.G.get('dmem'))
    self.G['dmem'] = dmem
    # x1 = LN1(t1), t1 = x + out_proj(sa), sa = mha(qkv(xb))
    dt1 = ops.layernorm_bwd(dx1, r['t1'], m1, r1, l.norm1.weight, g(l.norm1.weight), g(l.norm1.bias))
    sa = l.self_attn
    dz = ops.act_bwd(dt1, None, ACT_NONE, 1, rows, d, layout=2, dbias=g(sa.out_proj.bias), drop=dp[1])
    dsa = self.linear_bwd(dz, r['sa'], g(sa.out_proj.weight), packed(sa.out_proj.weight, 'linear_t'), d, d,
                          out_f32=False)
    qkv = r['qkv']
    dqkv = torch.empty_like(qkv)
    s3 = (nq * 3 * d, 3 * d)
    ops.small_mha_bwd(qkv, qkv, qkv, dsa, dqkv, dqkv, dqkv, b, heads, nq, nq, hd, s3, s3, s3, s3, s3, s3,
                      offs=(0, d, 2 * d, 0, d, 2 * d), drop=dp[0])
    ops.act_bwd(dqkv, None, ACT_NONE, 1, rows, 3 * d, dbias=g(sa.in_proj_bias), want_dz=False)
    dx = self.linear_bwd(dqkv, r['xb'], g(sa.in_proj_weight), packed(sa.in_proj_weight, 'linear_t'), 3 * d, d, res=dt1)
    self.G[id(r['x_in'])] = dx

  def planner_mem(self, r):
    from . import _lib  # pylint: disable=import-outside-toplevel
    st, m = self.st, self.eng.m
    g = st.g
    b, n_pix, n_mem, d = r['b'], r['n_pix'], r['n_mem'], r['d']
    dmem = self.G.pop('dmem')  # (B*n_mem, d) f32: accumulated over every decoder pass
    # extra sensor token = row n_pix of every sample
    ese, vn = m.extra_sensor_encoder, m.velocity_normalization
    training = r['training']
    _lib.check(_lib.load().tfpp_extra_sensor_token_bwd(
        r['ego_vel'].float().contiguous().data_ptr(), r['command'].float().contiguous().data_ptr(),
        0.0 if training else float(vn.running_mean[0]), 1.0 if training else float(vn.running_var[0]), int(training),
        ese[0].weight.data_ptr(), ese[0].bias.data_ptr(), ese[2].weight.data_ptr(), ese[2].bias.data_ptr(),
        dmem.data_ptr() + 4 * n_pix * d, n_mem * d, g(ese[0].weight).data_ptr(), g(ese[0].bias).data_ptr(),
        g(ese[2].weight).data_ptr(), g(ese[2].bias).data_ptr(), g(m.extra_sensor_pos_embed).data_ptr(), b,
        r['command'].shape[1], ese[0].weight.shape[0], d, ops._stream()), 'extra_sensor_bwd')  # pylint: disable=protected-access
    # pixel tokens: change_channel 1x1 conv (+ constant positional encoding)
    cc = m.change_channel
    fused = r['fused']
    cf = fused.shape[3]
    dz = ops.cast_rows(dmem, b, n_mem, 0, n_pix, d, dbias=g(cc.bias))
    dfused = self.linear_bwd(dz, fused.view(b * n_pix, cf), g(cc.weight).view(d, cf), packed(cc.weight, 'linear_t'), d,
                             cf, out_f32=False)
    self.add(fused, dfused.view(fused.shape))

  # -- driver --------------------------------------------------------------------------------------------------
  def run(self, tape, seeds):
    eng = self.eng
    for r in tape:
      if r['op'] == 'planner_mem':
        self.eng_mem = r['mem'].view(-1, r['d'])
    # seeds of the perspective decoders belong to the last conv_bias record of each decoder
    self._named_seeds = {}
    for key, dec in (('depth', 'depth_decoder'), ('semantic', 'semantic_decoder')):
      if key in seeds and hasattr(eng.m, dec):
        self._named_seeds[id(getattr(eng.m, dec).deconv3[2])] = seeds.pop(key)
    if not any(r.get('side') for r in tape):
      for r in reversed(tape):
        self._dispatch(r, seeds)
      return
    # Records tagged by Engine.forward replay on their side stream: the planner next to the dense heads, the LiDAR branch
    # of every stage next to the image branch.  A side segment starts after the main-stream work it depends on (the loss
    # seeds / the fusion block above it: an event recorded on the main stream at that point), and the main stream joins
    # a side stream before the first record that consumes its results.  Gradients popped while on a side stream are kept
    # alive until that join: they were allocated on the main stream, whose allocator would otherwise hand the block out
    # again while the side stream is still reading it.
    dev = self.st.grad.device
    main = torch.cuda.current_stream()
    streams = {k: eng.side_stream(dev, k) for k in {r['side'] for r in tape if r.get('side')}}
    keep = {k: [] for k in streams}
    keep['seeds'] = list(seeds.values())
    fork = torch.cuda.Event()
    fork.record(main)
    first_planner = min((i for i, r in enumerate(tape) if r.get('side') == 'planner'), default=-1)
    pending_join = set()
    outer = self

    class KeepDict(dict):
      def pop(self, *a):  # pylint: disable=arguments-differ
        v = dict.pop(self, *a)
        if outer._on_side is not None and torch.is_tensor(v):
          keep[outer._on_side].append(v)
        return v

    self.G = KeepDict(self.G)
    self._on_side = None

    def join(k):
      main.wait_stream(streams[k])
      keep[k].clear()
      pending_join.discard(k)

    prev_side = None
    for idx in range(len(tape) - 1, -1, -1):
      r = tape[idx]
      k = r.get('side')
      if k:
        if prev_side != k:
          streams[k].wait_event(fork)  # first record of a side segment: wait for the main-stream work above it
        self._on_side = k
        with torch.cuda.stream(streams[k]):
          self._dispatch(r, seeds)
        self._on_side = None
        pending_join.add(k)
      else:
        need = r.get('needs_side')
        if need and need in pending_join:
          join(need)
        if 'planner' in pending_join and idx < first_planner:
          join('planner')  # the backbone consumes the planner's gradient of the fused features
        self._dispatch(r, seeds)
        if need:
          fork = torch.cuda.Event()   # the LiDAR stage below this fusion block may start once it is done
          fork.record(main)
      prev_side = k
    for k in list(pending_join):
      join(k)
    self._on_side = None
    keep.clear()

  def _dispatch(self, r, seeds):
    op = r['op']
    if op == 'conv_bias':
      if id(r['conv']) in self._named_seeds:
        seeds[id(r['y'])] = self._named_seeds.pop(id(r['conv']))
      self.conv_bias(r, seeds)
    elif op == 'conv_bn':
      self.conv_bn(r)
    elif op == 'downsample':
      self.downsample(r)
    elif op == 'se':
      self.se(r)
    elif op == 'stem':
      self.stem(r)
    elif op == 'bilinear':
      self.bilinear(r)
    elif op == 'untokenise':
      self.untokenise(r)
    elif op == 'gpt_block':
      self.gpt_block(r)
    elif op == 'tokenise':
      self.tokenise(r)
    elif op == 'center_head':
      self.center_head(r, seeds)
    elif op == 'bev_tail':
      self.bev_tail(r, seeds)
    elif op == 'gru_cell_head':
      self.gru_cell_head(r, seeds)
    elif op == 'mlp_join':
      self.mlp_join(r)
    elif op == 'global_fuse':
      self.global_fuse(r)
    elif op == 'conv_in':
      self.conv_in(r)
    elif op == 'bev_lift':
      self.bev_lift(r)
    elif op == 'bev_stem':
      self.bev_stem(r)
    elif op == 'planner_head':
      self.planner_head(r, seeds)
    elif op == 'dec_layer':
      self.dec_layer(r)
    elif op == 'planner_mem':
      self.planner_mem(r)
    elif op == 'planner_queries':
      self.planner_queries(r)
    else:
      raise RuntimeError(f'no backward handler for {op}')


# ---------------------------------------------------------------------------------------------------------------
# trainer
# ---------------------------------------------------------------------------------------------------------------
def training_forward(eng, st, inputs):
  """Training-mode forward with a tape (list of saved activations) for training.Backward.  Returns (outputs, tape)."""
  eng.tape = []
  fused = st.counter_mask is not None
  eng.batch_counters_fused = fused
  eng.bn_seen = None if fused else []
  if fused:
    st.count_batches()
  try:
    out = eng.forward(inputs['rgb'], inputs['lidar_bev'], inputs['target_point'], inputs['ego_vel'], inputs['command'],
                      training=True)
    tape = eng.tape
  finally:
    eng.tape = None
    eng.batch_counters_fused = False
    if not fused and eng.bn_seen is not None and torch.cuda.is_available() and not torch.cuda.is_current_stream_capturing():
      st.learn_counter_mask(eng.bn_seen)
    eng.bn_seen = None
  return out, tape


def allreduce_flat(grad, group, bucket_elems):
  """Sum-all-reduce a flat gradient buffer in buckets (async, then wait): DDP's exchange step (train.py:516) without
  the per-parameter hooks.  Works on any backend (NCCL on the GPUs, gloo in the CPU tests)."""
  works = []
  for s in range(0, grad.numel(), bucket_elems):
    works.append(torch.distributed.all_reduce(grad[s:s + bucket_elems], group=group, async_op=True))
  for w in works:
    w.wait()


class Trainer:
  """One process per GPU.  step(batch) = forward + fused losses + backward + (bucketed all-reduce) + AdamW."""

  def __init__(self, model, lr=3e-4, weight_decay=0.01, loss_weights=None, process_group=None, bucket_mb=64,
               use_optim_groups=False, exchange=None):
    """exchange (world > 1): 'peer' = gradients reduce-scattered out of the peers' buffers over NVLink inside the fused
    AdamW kernel, parameters pushed back (csrc/peer_exchange.cu; the whole step stays one CUDA graph); 'nccl' = bucketed
    NCCL all-reduce between two graphs (round 1).  Default: 'peer' on CUDA (TFPP_EXCHANGE overrides), 'nccl' otherwise
    (the gloo CPU tests)."""
    self.model = model
    self.eng = model.engine
    if getattr(model, '_boundary', None) is not None:
      raise RuntimeError('this model already trains through the autograd boundary (model(...) in training mode)')
    object.__setattr__(model, '_trainer_owned', True)
    world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
    on_cuda = next(model.parameters()).is_cuda
    if exchange is None:
      exchange = os.environ.get('TFPP_EXCHANGE', 'peer' if on_cuda else 'nccl')
    if exchange not in ('peer', 'nccl'):
      raise ValueError(exchange)
    self.xchg = None
    self.local_only = False   # bench.py's rank-local instrumented step: skip the exchange
    alloc = None
    if world > 1 and exchange == 'peer':
      from . import peer  # pylint: disable=import-outside-toplevel

      def alloc(total):
        self.xchg = peer.PeerExchange(process_group, total)
        return self.xchg.param, self.xchg.grad
    self.st = FlatState(model, include_frozen=True, alloc=alloc)
    if use_optim_groups:  # train.py:522-525: decay / no-decay split by parameter name and module type
      groups = model.create_optimizer_groups(weight_decay)
      self.st.set_flags(no_decay=[p for g in groups if g['weight_decay'] == 0.0 for p in g['params']])
    self.lr, self.wd = lr, weight_decay
    self.keys = loss_keys(model.config)
    w = loss_weights or {k: 1.0 for k in self.keys}
    tot = sum(w.values())
    self.loss_weights = {k: v / tot for k, v in w.items()}  # train.py:452-456
    self.pg = process_group
    self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
    self.bucket_elems = bucket_mb * 1024 * 1024 // 4
    self.plan = eng_mod.PackPlan(self.st.flat) if self.st.flat.is_cuda else None
    if self.world > 1:
      # DistributedDataParallel's constructor broadcasts rank 0's parameters and buffers (train.py:516): replicas that
      # were initialised or loaded differently must not train diverged copies
      torch.distributed.broadcast(self.st.flat, src=torch.distributed.get_global_rank(process_group, 0), group=process_group)
      for buf in model.buffers():
        if buf.is_floating_point():
          torch.distributed.broadcast(buf, src=torch.distributed.get_global_rank(process_group, 0), group=process_group)

  def forward_backward(self, inputs, labels):
    eng, st = self.eng, self.st
    st.zero_grad()
    out, tape = training_forward(eng, st, inputs)
    losses, seeds = compute_losses(eng, out, labels, self.loss_weights, loss_bias_targets(eng, st))
    Backward(eng, st).run(tape, seeds)
    return out, losses

  def allreduce(self):
    """DDP semantics (train.py:516): sum-all-reduce the flat gradient in buckets, averaged by world size inside the
    AdamW kernel (grad_scale)."""
    if self.world == 1 or self.xchg is not None or self.local_only:
      return   # peer mode: the reduction happens inside the optimizer kernel
    allreduce_flat(self.st.grad, self.pg, self.bucket_elems)

  def _with_plan(self, fn):
    prev_plan = eng_mod._PLAN[0]  # pylint: disable=protected-access
    eng_mod._PLAN[0] = self.plan  # pylint: disable=protected-access
    try:
      return fn()
    finally:
      eng_mod._PLAN[0] = prev_plan  # pylint: disable=protected-access

  def optimizer_step(self, lr='default'):
    """AdamW over the flat buffers (gradient averaged over the ranks inside the kernel) + weight-pack refresh."""
    lr = self.lr if lr == 'default' else lr
    if self.xchg is not None and not self.local_only:
      st = self.st
      st.step_count += 1
      if lr is not None:
        st.dev_state[1:2].fill_(lr)
      self.xchg.step(st, weight_decay=self.wd)   # reduce-scatter + AdamW on the owned shard + parameter all-gather
      eng_mod.PARAM_EPOCH[0] += 1
    else:
      self.st.adamw_step(lr, weight_decay=self.wd, grad_scale=1.0 / (1 if self.local_only else self.world))
    if self.plan is not None:
      self.plan.refresh()  # one gather kernel: every bf16 weight pack follows the new parameters
      if not torch.cuda.is_current_stream_capturing():
        self.plan.finalize()  # adopt the packs first seen in this (eager) step

  def step(self, inputs, labels, lr='default'):
    def run():
      out, losses = self.forward_backward(inputs, labels)
      self.allreduce()
      self.optimizer_step(lr)
      return out, losses
    return self._with_plan(run)

  # ---- CUDA-graph replay of the step (removes ~2000 Python-issued launches per step from the critical path)
  def capture(self, inputs, labels, points=None, split=None):
    """Capture pillar scatter (if raw ``points`` are given) + step with static input buffers.  One rank: ONE CUDA graph
    for the whole step.  Data parallel (or ``split=True``): two graphs — forward/backward and optimizer — with the NCCL
    gradient all-reduce issued eagerly between the two replays (collectives stay out of the capture)."""
    self._sin = {k: v.clone() for k, v in inputs.items()}
    self._slab = {k: v.clone() for k, v in labels.items()}
    self._spts = points.clone() if points is not None else None
    self.st.dev_state[1:2].fill_(self.lr)
    self._split = (self.world > 1 and self.xchg is None) if split is None else bool(split)

    def fwd_bwd():
      if self._spts is not None:
        self._sin['lidar_bev'] = ops.pillar_scatter(self._spts, use_ground_plane=bool(self.eng.cfg.use_ground_plane))
      return self._with_plan(lambda: self.forward_backward(self._sin, self._slab))

    def opt():
      self._with_plan(lambda: self.optimizer_step(lr=None))

    def body():
      r = fwd_bwd()
      self.allreduce()
      opt()
      return r

    # the two warm-up steps are real optimizer steps: snapshot everything they mutate and put it back afterwards, so
    # that capture() has no side effect on the training state
    st = self.st
    snap = [(t, t.clone()) for t in (st.flat, st.exp_avg, st.exp_avg_sq, st.max_exp_avg_sq, st.dev_state,
                                     st.batch_counters)]
    snap += [(b, b.clone()) for b in self.model.buffers() if b.is_floating_point()]
    step_count = st.step_count
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
      for _ in range(2):
        body()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    with torch.no_grad():
      for t, saved in snap:
        t.copy_(saved)
    st.step_count = step_count
    st.dev_state[1:2].fill_(self.lr)
    if self.plan is not None:
      self.plan.refresh()
    eng_mod.PARAM_EPOCH[0] += 1
    torch.cuda.synchronize()
    from . import _lib  # pylint: disable=import-outside-toplevel
    _lib.reset_launch_count()
    self.graph = torch.cuda.CUDAGraph()
    self.graph_opt = None
    if not self._split:
      with torch.cuda.graph(self.graph):
        out, losses = body()
        self._gloss = torch.stack([losses[k] for k in self.keys])
    else:
      with torch.cuda.graph(self.graph):
        out, losses = fwd_bwd()
        self._gloss = torch.stack([losses[k] for k in self.keys])
      self.graph_opt = torch.cuda.CUDAGraph()
      with torch.cuda.graph(self.graph_opt):
        opt()
    self._gout = out
    self.launches_per_step = _lib.launch_count()  # libtfpp kernels recorded into the graph(s)
    return self

  # ---- input prefetch (opt-in; not yet run on a GPU): H2D copies of the NEXT step on a copy stream, overlapping the
  # current replay; replay_staged() hands them to the graph's static buffers with device-to-device copies
  def stage(self, inputs=None, labels=None, points=None):
    """Start copying the next step's (pinned host) tensors into a second set of device buffers on a side stream."""
    if getattr(self, '_copy_stream', None) is None:
      self._copy_stream = torch.cuda.Stream()
      self._stg_in = {k: torch.empty_like(v) for k, v in self._sin.items()}
      self._stg_lab = {k: torch.empty_like(v) for k, v in self._slab.items()}
      self._stg_pts = torch.empty_like(self._spts) if self._spts is not None else None
      self._stg_ready = torch.cuda.Event()
      self._stg_free = torch.cuda.Event()
      self._stg_free.record(torch.cuda.current_stream())
    cs = self._copy_stream
    cs.wait_event(self._stg_free)  # the previous hand-over has read the staging buffers
    with torch.cuda.stream(cs):
      if inputs is not None:
        for k, v in inputs.items():
          if k in self._stg_in and not (k == 'lidar_bev' and self._spts is not None):
            self._stg_in[k].copy_(v, non_blocking=True)
      if labels is not None:
        for k, v in labels.items():
          self._stg_lab[k].copy_(v, non_blocking=True)
      if points is not None:
        self._stg_pts.copy_(points, non_blocking=True)
      self._stg_ready.record(cs)
    self._stg_what = (inputs is not None, labels is not None, points is not None)

  def replay_staged(self):
    """Replay the step on the tensors passed to the last stage() call."""
    main = torch.cuda.current_stream()
    main.wait_event(self._stg_ready)
    has_in, has_lab, has_pts = self._stg_what
    if has_in:
      for k, v in self._stg_in.items():
        if not (k == 'lidar_bev' and self._spts is not None):
          self._sin[k].copy_(v, non_blocking=True)
    if has_lab:
      for k, v in self._stg_lab.items():
        self._slab[k].copy_(v, non_blocking=True)
    if has_pts:
      self._spts.copy_(self._stg_pts, non_blocking=True)
    self._stg_free.record(main)
    return self.replay()

  def replay(self, inputs=None, labels=None, points=None):
    """Copy new data into the static buffers (non-blocking; pinned host tensors welcome) and replay the step."""
    if inputs is not None:
      for k, v in inputs.items():
        if k in self._sin and not (k == 'lidar_bev' and self._spts is not None):
          self._sin[k].copy_(v, non_blocking=True)
    if labels is not None:
      for k, v in labels.items():
        self._slab[k].copy_(v, non_blocking=True)
    if points is not None:
      self._spts.copy_(points, non_blocking=True)
    self.graph.replay()
    if self.graph_opt is not None:
      self.allreduce()
      self.graph_opt.replay()
    # the graph rewrote parameters and BatchNorm running statistics through raw pointers (no version counter moved):
    # every cached weight pack / folded BatchNorm affine outside the PackPlan is stale now
    eng_mod.PARAM_EPOCH[0] += 1
    return self._gout, self._gloss
