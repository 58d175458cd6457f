"""Autograd-compatible training boundary: the reference's own training loop (team_code/train.py:776-820,883-916) calls

    pred = model(rgb=..., lidar_bev=..., target_point=..., ego_vel=..., command=...)      # train.py:776-780
    losses = model.compute_loss(**preds, **labels)                                        # train.py:797-820
    loss = sum(w_k * losses[k]); loss.backward(); optimizer.step()                        # train.py:889-908

on a (DistributedDataParallel-wrapped) LidarCenterNet.  SURVEY.md §8(b): outputs must be ordinary autograd-tracked
tensors.  This module makes the B200 engine look like that to torch.autograd without putting torch kernels on the
path:

* ``_Step`` — ONE autograd.Function for the whole forward.  Its inputs are the five input tensors and every trainable
  parameter (so DistributedDataParallel's AccumulateGrad hooks fire); its outputs are the six dense prediction
  buffers; forward records the engine's tape, backward replays it (training.Backward) and hands back views of a flat
  fp32 gradient buffer.
* ``_Loss`` — the fused loss kernels (csrc/loss.cu) as an autograd.Function returning the 10 loss values.  Its backward
  runs the loss kernels again with d(total)/d(loss_k) read ON THE DEVICE and parks the seed gradients (already in the
  layout the backward GEMMs consume) on the boundary object; towards autograd it returns zero-stride placeholders that
  ``_Step.backward`` recognises.  Predictions that reach ``_Step.backward`` with ordinary gradients (a user's own torch
  loss on the outputs) are converted by tfpp_act_bwd (NCHW f32 -> NHWC bf16 seed + bias gradient) instead.

Parameters live in training.FlatState's flat buffer (state_dict-compatible views), so a torch optimizer updates them
in place; the bf16 weight packs follow through one tfpp_gather_pack launch at the start of the next forward.
"""
import torch

from . import engine as eng_mod
from . import ops
from . import training
from .ops import ACT_NONE, ACT_SIGMOID, BF16, F32
from .training import loss_keys


class TrainBoundary:
  """Per-model state of the autograd path (created lazily by LidarCenterNet.forward in training mode)."""

  def __init__(self, model):
    self.model = model
    self.eng = model.engine
    self.st = training.FlatState(model, assign_grad=False, include_frozen=True)
    self.gbufs = [self.st.grad, torch.zeros_like(self.st.grad)]
    self.plan = eng_mod.PackPlan(self.st.flat)
    self.zero = torch.zeros((), dtype=F32, device=self.st.flat.device)  # storage of the placeholder gradients
    self.stash = None   # seeds + bias gradients parked by _Loss.backward for the next _Step.backward
    self._sig = None

  # ------------------------------------------------------------------------------------------------ helpers
  def placeholder(self, like):
    return self.zero.expand(like.shape)

  def is_placeholder(self, g):
    return g is not None and g.untyped_storage().data_ptr() == self.zero.untyped_storage().data_ptr()

  def with_plan(self, fn):
    prev = eng_mod._PLAN[0]  # pylint: disable=protected-access
    eng_mod._PLAN[0] = self.plan  # pylint: disable=protected-access
    try:
      return fn()
    finally:
      eng_mod._PLAN[0] = prev  # pylint: disable=protected-access

  def sync_packs(self):
    """The parameters may have been rewritten by torch ops (optimizer.step(), load_state_dict) since the last forward:
    one gather rebuilds every plan-owned weight pack."""
    sig = sum(p._version for p in self.st.params)  # pylint: disable=protected-access
    if sig != self._sig:
      self.plan.refresh_all()
      self._sig = sig

  def pick_grad_buffer(self):
    """Write into the buffer the live .grad tensors do NOT alias: AccumulateGrad adds our result onto an existing
    .grad in place (gradient accumulation over micro-batches) and adopts it when .grad is None."""
    lo = self.gbufs[0].data_ptr()
    hi = lo + self.gbufs[0].numel() * 4
    for p in self.st.params:
      if p.grad is not None:
        ptr = p.grad.data_ptr()
        self.st.grad = self.gbufs[1] if lo <= ptr < hi else self.gbufs[0]
        return
    self.st.grad = self.gbufs[0]

  # ------------------------------------------------------------------------------------------------ forward
  def forward(self, rgb, lidar_bev, target_point, ego_vel, command):
    ps = tuple(p for p in self.st.params if p.requires_grad)
    outs = _Step.apply(self, rgb, lidar_bev, target_point, ego_vel, command, *ps)
    cfg = self.model.config
    it = iter(outs)
    ts, cp = (next(it), next(it)) if cfg.use_controller_input_prediction else (None, None)
    sem, bev, depth, maps = next(it), next(it), next(it), next(it)
    wp = next(it) if cfg.use_wp_gru else None
    head = self.model.head
    sizes = [getattr(head, n)[2].weight.shape[0] for n in head.head_names()]
    views, o = [], 0
    for s in sizes:
      views.append(maps[:, o:o + s])
      o += s
    bb = (views[0], views[1], views[2], views[3], views[4], None, None)
    return (wp, ts, cp, sem, bev, depth.squeeze(1), bb, None, None, None)

  def seeds_from(self, tape_out, gouts):
    """Seed gradients for training.Backward from what autograd delivered (+ what _Loss.backward parked)."""
    st, m = self.st, self.model
    ts, cp, sem, bev, depth, maps, wp = tape_out
    g_ts, g_cp, g_sem, g_bev, g_depth, g_maps, g_wp = gouts
    stash, self.stash = self.stash, None
    seeds = dict(stash['seeds']) if stash is not None else {}
    if stash is not None:  # bias gradients of the heads' last convs, computed by the loss kernels
      tgt = training.loss_bias_targets(self.eng, st)
      for k, v in stash['bias'].items():
        tgt[k].add_(v)
    real = lambda g: g is not None and not self.is_placeholder(g)
    b = sem.shape[0]
    if wp is not None and (real(g_wp) or 'planner_wp' not in seeds):
      d0 = seeds.get('planner_wp', (None, None))[0]
      dwp = g_wp.contiguous().float() if real(g_wp) else torch.zeros_like(wp)
      seeds['planner_wp'] = (dwp + d0 if d0 is not None else dwp, None)
    if ts is not None and (real(g_ts) or real(g_cp) or 'planner' not in seeds):
      dcp0, dl0 = seeds.get('planner', (None, None))
      dcp = g_cp.contiguous().float() if real(g_cp) else torch.zeros_like(cp)
      dl = g_ts.contiguous().float() if real(g_ts) else torch.zeros_like(ts)
      seeds['planner'] = (dcp + dcp0 if dcp0 is not None else dcp, dl + dl0 if dl0 is not None else dl)

    def dense(key, g, y, act, n_limit, cpad, bias):
      if real(g) or key not in seeds:
        c = y.shape[1]
        hw = y.shape[2] * y.shape[3]
        if real(g):
          dz = ops.act_bwd(g.contiguous().float(), y if act != ACT_NONE else None, act, b, hw, c, layout=1,
                           act_n_limit=n_limit, dbias=bias, channels_padded=cpad).view(b, y.shape[2], y.shape[3], cpad)
        else:
          dz = torch.zeros((b, y.shape[2], y.shape[3], cpad), dtype=ops.act_dtype(), device=y.device)
        if key in seeds:
          ops.add_bf16(dz, seeds[key], out=dz)
        seeds[key] = dz

    tgt = training.loss_bias_targets(self.eng, st)
    dense('semantic', g_sem, sem, ACT_NONE, 0, 16, tgt['semantic'])
    dense('depth', g_depth, depth, ACT_SIGMOID, 0, 16, tgt['depth'])
    dense('center', g_maps, maps, ACT_SIGMOID, m.config.num_bb_classes, 24, tgt['center'])
    if real(g_bev):
      d = g_bev.contiguous().float()
      seeds['bev'] = d + seeds['bev'] if 'bev' in seeds else d
    elif 'bev' not in seeds:
      seeds['bev'] = torch.zeros_like(bev)
    return seeds


class _Step(torch.autograd.Function):
  """LidarCenterNet.forward (model.py:279-392) + its whole backward as one autograd node."""

  @staticmethod
  def forward(ctx, bnd, rgb, lidar_bev, target_point, ego_vel, command, *params):
    del params  # the engine reads them through the module tree; they are inputs so that autograd routes their gradients
    eng, st = bnd.eng, bnd.st
    bnd.sync_packs()
    out, tape = bnd.with_plan(lambda: training.training_forward(eng, st, dict(
        rgb=rgb, lidar_bev=lidar_bev, target_point=target_point, ego_vel=ego_vel, command=command)))
    bnd.plan.finalize()
    wp, ts, cp, sem, bev = out[0], out[1], out[2], out[3], out[4]
    depth = training._base_of(out[5])  # (B,1,H,W)  pylint: disable=protected-access
    maps = training._base_of(out[6][0])  # (B,21,64,64)  pylint: disable=protected-access
    ctx.bnd, ctx.tape, ctx.outs = bnd, tape, (ts, cp, sem, bev, depth, maps, wp)
    return tuple(t for t in (ts, cp, sem, bev, depth, maps, wp) if t is not None)

  @staticmethod
  @torch.autograd.function.once_differentiable
  def backward(ctx, *gouts):
    bnd = ctx.bnd
    eng, st = bnd.eng, bnd.st
    bnd.pick_grad_buffer()
    st.zero_grad()
    eng.new_arena(st.flat.device)
    it = iter(gouts)
    full = tuple(next(it) if t is not None else None for t in ctx.outs)   # gradients in (ts, cp, sem, bev, depth, maps, wp) slots
    seeds = bnd.seeds_from(ctx.outs, full)
    # the tape is read-only for Backward: it stays on ctx (freed with the autograd graph) so that
    # loss.backward(retain_graph=True) followed by a second backward works like it does for torch modules
    bnd.with_plan(lambda: training.Backward(eng, st).run(ctx.tape, seeds))
    grads = tuple(st.g(p) for p in st.params if p.requires_grad)
    return (None,) * 6 + grads


class _Loss(torch.autograd.Function):
  """compute_loss (model.py:394-445, center_net.py:77-123) on the fused loss kernels; returns the 10 values in
  training.LOSS_KEYS order as one (10,) tensor."""

  @staticmethod
  def forward(ctx, eng, bnd, labels, present, *preds):
    it = iter(preds)
    ts, cp, sem, bev, depth, maps, wp = (next(it) if f else None for f in present)
    losses, _ = training.compute_losses(eng, (wp, ts, cp, sem, bev, depth, maps), labels, want_seeds=False)
    ctx.eng, ctx.bnd, ctx.labels, ctx.present = eng, bnd, labels, present
    ctx.save_for_backward(*preds)
    return torch.stack([losses[k] for k in loss_keys(eng.cfg)])

  @staticmethod
  @torch.autograd.function.once_differentiable
  def backward(ctx, gvals):
    eng, bnd = ctx.eng, ctx.bnd
    if bnd is None:
      raise RuntimeError('compute_loss().backward() needs predictions produced by this model in training mode')
    it = iter(ctx.saved_tensors)
    ts, cp, sem, bev, depth, maps, wp = (next(it) if f else None for f in ctx.present)
    dev = sem.device
    bias = {'semantic': torch.zeros(sem.shape[1], dtype=F32, device=dev), 'depth': torch.zeros(1, dtype=F32, device=dev),
            'center': torch.zeros(maps.shape[1], dtype=F32, device=dev)}
    eng.new_arena(dev, 64)
    _, seeds = training.compute_losses(eng, (wp, ts, cp, sem, bev, depth, maps), ctx.labels, bias_grads=bias,
                                       w_dev=gvals.contiguous().float())
    bnd.stash = {'seeds': seeds, 'bias': bias}
    return (None, None, None, None) + tuple(bnd.placeholder(t) for t in ctx.saved_tensors)


def compute_loss(model, pred_wp, waypoint_label, pred_target_speed, pred_checkpoint, pred_semantic, pred_bev_semantic, pred_depth,
                 pred_bounding_box, target_speed_label, checkpoint_label, semantic_label, bev_semantic_label, depth_label,
                 center_heatmap_label, wh_label, yaw_class_label, yaw_res_label, offset_label, pixel_weight_label,
                 avg_factor_label):
  """Drop-in body of LidarCenterNet.compute_loss: dict of the 10 loss tensors (model.py:394-445)."""
  eng = model.engine
  dev = pred_semantic.device
  lab = {}
  if pred_wp is not None:
    lab['waypoint'] = waypoint_label.to(dev, F32).contiguous()
  if pred_target_speed is not None:
    lab.update(target_speed=target_speed_label.to(dev, torch.long).contiguous(),
               checkpoint=checkpoint_label.to(dev, F32).contiguous())
  lab.update({
      'semantic': semantic_label.to(dev, torch.long).contiguous(),
      'bev_semantic': bev_semantic_label.to(dev, torch.long).contiguous(),
      'depth': depth_label.to(dev, F32).contiguous(),
      'center_heatmap': center_heatmap_label.to(dev, F32).contiguous(),
      'wh': wh_label.to(dev, F32).contiguous(),
      'offset': offset_label.to(dev, F32).contiguous(),
      'yaw_class': yaw_class_label.to(dev, torch.long).contiguous(),
      'yaw_res': yaw_res_label.to(dev, F32).contiguous(),
      'pixel_weight': pixel_weight_label.to(dev, F32).contiguous(),
      'avg_factor': avg_factor_label.to(dev, F32).contiguous(),
  })
  bb = pred_bounding_box
  base = bb[0]._base  # pylint: disable=protected-access
  n_maps = sum(t.shape[1] for t in bb[:5])
  if base is not None and base.dim() == 4 and base.shape[1] == n_maps and base.is_contiguous() and \
      bb[0].data_ptr() == base.data_ptr():
    maps = base  # the fused (B,21,64,64) buffer the five views were cut from
  else:
    maps = torch.cat(bb[:5], dim=1)
  depth = pred_depth if pred_depth.dim() == 4 else pred_depth.unsqueeze(1)
  bnd = getattr(model, '_boundary', None)
  eng.new_arena(dev, 64)
  slots = (pred_target_speed, pred_checkpoint, pred_semantic, pred_bev_semantic, depth, maps, pred_wp)
  present = tuple(t is not None for t in slots)
  vals = _Loss.apply(eng, bnd, lab, present, *[t.contiguous() for t in slots if t is not None])
  order = ('loss_wp',) + training.LOSS_KEYS   # the reference's dict order (model.py:399-443)
  keys = loss_keys(eng.cfg)
  return {k: vals[keys.index(k)] for k in order if k in keys}
