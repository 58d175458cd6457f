"""GPU tests of the autograd-compatible training boundary (carla_garage_b200/boundary.py): the reference's own train
loop — model(**inputs) -> model.compute_loss(...) -> weighted sum -> loss.backward() -> torch optimizer
(team_code/train.py:776-820,883-916) — must run unmodified on the B200 engine and produce the gradients of the fused
Trainer path."""
import os
import re

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def rel(a, b):
  a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
  return float((a - b).norm() / (b.norm() + 1e-30))


ZERO_GRAD = re.compile(r'(attn\.key\.bias|transformers\.[012]\.ln_f\.bias|img_channel_to_lidar\.[012]\.bias)$')


def zero_grad_param(name):
  """Parameters whose gradient is exactly zero in exact arithmetic, i.e. pure rounding noise on both sides of any
  comparison: the key bias (softmax shift invariance) and the per-channel constants that the first three fusion stages
  add to the feature maps right before a training-mode BatchNorm removes every per-channel constant again (ln_f.bias
  reaches the maps through the bilinear up-sampling / the 1x1 back-projection, whose bias is the same kind of term)."""
  return ZERO_GRAD.search(name) is not None


def grad_scale(grads):
  """rms over every element of a dict of gradients: the yardstick for parameters whose own gradient is (analytically)
  zero or nearly so — a per-channel constant added before a training-mode BatchNorm (attn.key.bias, ln_f.bias and the
  residual-stream biases of the first three fusion GPTs) — where a relative error only measures rounding noise."""
  tot = sum(float(g.double().pow(2).sum()) for g in grads.values())
  cnt = sum(g.numel() for g in grads.values())
  return (tot / max(cnt, 1)) ** 0.5


def err(a, b, scale):
  """|a - b| relative to |b| + the gradient yardstick (see grad_scale)."""
  a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
  return float((a - b).norm() / (b.norm() + scale * b.numel() ** 0.5 + 1e-30))


def cos(a, b):
  a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
  return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))

# Two separate forward passes of this randomly initialised network differ by ~1 % at the outputs: the fp32 atomics of
# the BatchNorm / SE reductions sum in a different order, one flipped bf16 rounding is amplified ~1.2x per residual
# block (DESIGN.md "Numerics").  Comparisons ACROSS forwards are therefore loose (RUN2RUN); the tight comparisons below
# are made on ONE forward (two backward passes through the same autograd graph).
RUN2RUN = 5e-2


def _model(oracle_state):
  from carla_garage_b200.config import GlobalConfig
  from carla_garage_b200.nn import LidarCenterNet
  m = LidarCenterNet(GlobalConfig())
  m.load_state_dict(oracle_state, strict=True)
  return m.cuda().train()


def _data():
  from carla_garage_b200 import synth
  inp = {k: v.cuda() for k, v in synth.make_inputs(2, seed=11).items()}
  lab = {k: v.cuda().contiguous() for k, v in synth.make_labels(2, seed=13).items()}
  return inp, lab


def _reference_style_losses(model, out, lab):
  """The call train.py:797-820 makes (keyword arguments, None for the disabled heads)."""
  return model.compute_loss(pred_wp=out[0], pred_target_speed=out[1], pred_checkpoint=out[2], pred_semantic=out[3],
                            pred_bev_semantic=out[4], pred_depth=out[5], pred_bounding_box=out[6], waypoint_label=None,
                            target_speed_label=lab['target_speed'], checkpoint_label=lab['checkpoint'],
                            semantic_label=lab['semantic'], bev_semantic_label=lab['bev_semantic'],
                            depth_label=lab['depth'], center_heatmap_label=lab['center_heatmap'], wh_label=lab['wh'],
                            yaw_class_label=lab['yaw_class'], yaw_res_label=lab['yaw_res'], offset_label=lab['offset'],
                            velocity_label=None, brake_target_label=None, pixel_weight_label=lab['pixel_weight'],
                            avg_factor_label=lab['avg_factor'], pred_wp_1=out[8], selected_path=out[9])


def _torch_losses(model, out, lab):
  """model.py:394-445 + center_net.py:77-123 restated with plain torch ops on the GPU tensors (general autograd path)."""
  bb = out[6]
  loss = {}
  loss['loss_target_speed'] = F.cross_entropy(out[1], lab['target_speed'], weight=model.loss_speed.weight)
  loss['loss_checkpoint'] = torch.mean(torch.abs(out[2] - lab['checkpoint']))
  loss['loss_semantic'] = F.cross_entropy(out[3], lab['semantic'])
  valid = model.valid_bev_pixels.squeeze(1).int()
  vis = (valid - 1) + valid * lab['bev_semantic']
  loss['loss_bev_semantic'] = F.cross_entropy(out[4], vis.long(), ignore_index=-1)
  loss['loss_depth'] = F.l1_loss(out[5], lab['depth'])
  avg = lab['avg_factor'].sum() + torch.finfo(torch.float32).eps
  pw = lab['pixel_weight']
  p, t = bb[0], lab['center_heatmap']
  eps = 1e-12
  pos = t.eq(1).float()
  focal = -(p + eps).log() * (1 - p).pow(2) * pos - (1 - p + eps).log() * p.pow(2) * (1 - t).pow(4)
  loss['loss_center_heatmap'] = focal.sum() / avg
  loss['loss_wh'] = (torch.abs(bb[1] - lab['wh']) * pw).sum() / (avg * 2)
  loss['loss_offset'] = (torch.abs(bb[2] - lab['offset']) * pw).sum() / (avg * 2)
  loss['loss_yaw_class'] = (F.cross_entropy(bb[3], lab['yaw_class'], reduction='none') * pw[:, 0]).sum() / avg
  loss['loss_yaw_res'] = (F.smooth_l1_loss(bb[4], lab['yaw_res'], reduction='none') * pw[:, 0:1]).sum() / avg
  return loss


@pytest.fixture(scope='module')
def trainer_grads(oracle_state):
  """Gradients and losses of the fused Trainer path on the shared batch (the thing the autograd path must reproduce)."""
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  from carla_garage_b200.training import Trainer
  m = _model(oracle_state)
  tr = Trainer(m)
  inp, lab = _data()
  _, losses = tr.forward_backward(inp, lab)
  torch.cuda.synchronize()
  grads = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.requires_grad}
  return {k: float(v) for k, v in losses.items()}, grads


@pytest.mark.noisy
def test_reference_train_loop_runs_unmodified(oracle_state, trainer_grads):
  want_losses, want_grads = trainer_grads
  m = _model(oracle_state)
  opt = torch.optim.AdamW(m.parameters(), lr=3e-4, amsgrad=True)   # train.py:531, created before the first forward
  inp, lab = _data()
  m.eval()
  with torch.no_grad():
    eval_before = [t.clone() if torch.is_tensor(t) else t for t in m(**inp)]   # also warms the eval-mode weight-pack / BN-fold caches
  m.train()
  opt.zero_grad(set_to_none=False)                                  # train.py:880
  out = m(**inp)                                                   # train.py:776-780 (keyword call)
  assert out[1].grad_fn is not None and out[3].grad_fn is not None and out[6][0].grad_fn is not None
  assert out[0] is None and out[7] is None and out[6][5] is None
  assert out[5].shape == (2, 256, 1024) and out[6][3].shape == (2, 12, 64, 64)
  losses = _reference_style_losses(m, out, lab)
  assert set(losses) == set(want_losses)
  loss = torch.zeros(1, dtype=torch.float32, device='cuda')
  for k, v in losses.items():                                      # train.py:889-896
    loss += 0.1 * v
    assert abs(float(v.item()) - want_losses[k]) <= RUN2RUN * max(abs(want_losses[k]), 0.05), (k, float(v), want_losses[k])
  loss.backward()                                                  # train.py:898
  torch.cuda.synchronize()
  worst = 0.0
  scale1 = grad_scale(want_grads)
  for n, p in m.named_parameters():
    if not p.requires_grad:
      continue
    assert p.grad is not None, n
    if zero_grad_param(n):
      continue  # exactly zero in exact arithmetic (softmax shift invariance): pure rounding noise on both sides
    if float(want_grads[n].norm()) < 0.05 * scale1 * want_grads[n].numel() ** 0.5:
      continue   # (nearly) zero gradient: rounding noise on both sides
    c = cos(p.grad, want_grads[n])
    worst = min(worst if worst else 1.0, c)
    # same kernels, another forward pass (see RUN2RUN): the gradients of the first blocks' small parameters (BN bias,
    # squeeze-excite MLP with a handful of active units) are the noisiest, measured down to 0.59 between two runs
    assert c > 0.2, (n, c)   # measured 0.34 (s1.b2.se.fc1.weight: 6 active hidden units)
    assert 0.25 < float(p.grad.norm()) / float(want_grads[n].norm() + 1e-30) < 4.0, n
  names = [n for n, p in m.named_parameters() if p.requires_grad and not zero_grad_param(n)]
  whole = cos(torch.cat([dict(m.named_parameters())[n].grad.flatten() for n in names]),
              torch.cat([want_grads[n].flatten() for n in names]))
  assert whole > 0.9, whole   # the full gradient vector: a mis-routed parameter gradient would show up here
  print(f'  autograd path vs fused Trainer path (separate forwards): worst gradient cosine {worst:.4f}')
  before = {n: p.detach().clone() for n, p in list(m.named_parameters())[:8]}
  opt.step()                                                       # train.py:908
  opt.zero_grad(set_to_none=True)                                  # train.py:910
  assert any(float((p.detach() - before[n]).abs().max()) > 0 for n, p in list(m.named_parameters())[:8])
  # second iteration: the weight packs must follow the torch optimizer (one gather at the start of the forward)
  out2 = m(**inp)
  l2 = sum(0.1 * v for v in _reference_style_losses(m, out2, lab).values())
  l2.backward()
  opt.step()
  out3 = m(**inp)
  l3 = sum(0.1 * v for v in _reference_style_losses(m, out3, lab).values())
  assert float(l3) < float(loss), (float(loss), float(l2), float(l3))  # same batch three times: the loss goes down
  # eval after training steps equals a FRESH model loaded from the trained state_dict (no stale weight packs / BN folds)
  m.eval()
  with torch.no_grad():
    got = [t.clone() if torch.is_tensor(t) else t for t in m(**inp)]
    again = m(**inp)
  fresh = _model({k: v.detach().cpu() for k, v in m.state_dict().items()}).eval()
  with torch.no_grad():
    want = fresh(**inp)
  for i in (1, 2, 3, 4, 5):
    # two eval forwards of the same weights differ by the squeeze-excite atomics (RUN2RUN); stale packs / BatchNorm folds
    # would leave `got` at the pre-training outputs instead, several noise levels away
    noise = rel(again[i], got[i])
    assert rel(got[i], want[i]) < max(RUN2RUN, 4 * noise), (i, rel(got[i], want[i]), noise)
    assert rel(got[i], eval_before[i]) > 2 * rel(got[i], want[i]), (i, rel(got[i], eval_before[i]), rel(got[i], want[i]))


@pytest.mark.noisy
def test_general_autograd_path_torch_losses(oracle_state):
  """A user's own torch loss on the outputs (no fused loss kernels): ordinary gradients arrive at the boundary and are
  converted into seeds by tfpp_act_bwd.  Compared on ONE forward with the fused-loss path (two backward passes through
  the same graph), so the only difference is one more bf16 rounding of the seeds."""
  m = _model(oracle_state)
  inp, lab = _data()
  out = m(**inp)
  fused = _reference_style_losses(m, out, lab)
  plain = _torch_losses(m, out, lab)
  for k in fused:
    assert abs(float(fused[k]) - float(plain[k])) <= 2e-5 * max(1.0, abs(float(plain[k]))), k
  sum(0.1 * v for v in fused.values()).backward(retain_graph=True)
  torch.cuda.synchronize()
  ga = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.requires_grad}
  m.zero_grad(set_to_none=True)
  sum(0.1 * v for v in plain.values()).backward()
  torch.cuda.synchronize()
  worst = ('', 0.0)
  scale2 = grad_scale(ga)
  for n, p in m.named_parameters():
    if not p.requires_grad or zero_grad_param(n):
      continue
    e = err(p.grad, ga[n], scale2)
    if e > worst[1]:
      worst = (n, e)
    assert e < 6e-2, (n, e)   # measured 3.4e-2 (a token-summed bias gradient with heavy cancellation), 2.06e-2 on the first stage's squeeze-excite fc1 (18 hidden units, 6 active)
  print(f'  general vs fused seeds on one forward: worst parameter-gradient rel err {worst[1]:.2e} ({worst[0]})')


@pytest.mark.noisy
def test_gradient_accumulation_and_partial_losses(oracle_state):
  """Two backward passes without zero_grad accumulate (AccumulateGrad adds in place, the boundary alternates its flat
  buffers); a loss that touches only some outputs still yields a gradient for every parameter (zeros where unused)."""
  m = _model(oracle_state)
  inp, lab = _data()
  out = m(**inp)
  sum(0.1 * v for v in _reference_style_losses(m, out, lab).values()).backward()
  g1 = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.requires_grad}
  out = m(**inp)
  sum(0.1 * v for v in _reference_style_losses(m, out, lab).values()).backward()
  scale3 = grad_scale(g1)
  for n, p in m.named_parameters():
    if p.requires_grad and not zero_grad_param(n) and float(g1[n].norm()) > 0.05 * scale3 * g1[n].numel() ** 0.5:
      # another forward (RUN2RUN): loose per parameter, see test_reference_train_loop_runs_unmodified
      assert cos(p.grad, g1[n]) > 0.5 and 1.3 < float(p.grad.norm()) / float(g1[n].norm()) < 3.0, n
  # exact accumulation semantics on one forward: backward twice through the same graph doubles .grad
  m.zero_grad(set_to_none=True)
  out = m(**inp)
  tot = sum(0.1 * v for v in _reference_style_losses(m, out, lab).values())
  tot.backward(retain_graph=True)
  g1 = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.requires_grad}
  tot.backward()
  for n, p in m.named_parameters():
    if p.requires_grad and not zero_grad_param(n) and float(g1[n].norm()) > 0:
      assert err(p.grad, 2 * g1[n], 2 * grad_scale(g1)) < 6e-2, n   # two backward passes: fp32 atomics sum in another order (measured 1.2e-2)
  m.zero_grad(set_to_none=True)
  out = m(**inp)
  out[2].abs().mean().backward()  # checkpoints only
  p = dict(m.named_parameters())
  assert float(p['checkpoint_decoder.decoder.weight'].grad.abs().max()) > 0
  assert float(p['semantic_decoder.deconv3.2.weight'].grad.abs().max()) == 0
  assert float(p['backbone.image_encoder.stem.conv.weight'].grad.abs().max()) > 0


def test_head_loss_matches_compute_loss(oracle_state):
  m = _model(oracle_state).eval()
  inp, lab = _data()
  with torch.no_grad():
    out = m(**inp)
    full = _reference_style_losses(m, out, lab)
    bb = out[6]
    part = m.head.loss(bb[0], bb[1], bb[2], bb[3], bb[4], None, None, lab['center_heatmap'], lab['wh'],
                       lab['yaw_class'], lab['yaw_res'], lab['offset'], None, None, lab['pixel_weight'],
                       lab['avg_factor'])
    ref = _torch_losses(m, out, lab)
  for k, v in part.items():
    assert abs(float(v) - float(full[k])) <= 1e-5 * max(1.0, abs(float(full[k])))
  for k, v in full.items():  # fused loss kernels vs plain torch on identical predictions
    assert abs(float(v) - float(ref[k])) <= 2e-5 * max(1.0, abs(float(ref[k]))), (k, float(v), float(ref[k]))
