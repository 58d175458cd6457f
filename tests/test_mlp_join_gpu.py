"""GPU parity of the original TransFuser planner (config.transformer_decoder_join = False [+ use_wp_gru]):
global-pooled features + extra-sensor embedding -> MLP join -> autoregressive GRUCell heads
(GRUWaypointsPredictorTransFuser, model.py:870-913) + target-speed MLP; goldens from the unmodified reference
(tests/golden/make_golden_mlp_join.py)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def rel(a, b):
  a, b = torch.as_tensor(a).double().cpu().flatten(), torch.as_tensor(b).double().cpu().flatten()
  return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope='module')
def ops():
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  from carla_garage_b200 import ops as o
  return o


def _cfg():
  from carla_garage_b200.config import GlobalConfig
  cfg = GlobalConfig()
  cfg.transformer_decoder_join = False
  cfg.use_wp_gru = True
  return cfg


@pytest.mark.parametrize('b,steps,hidden,learn_origin,with_ts', [(3, 10, 64, True, True), (2, 8, 64, True, False),
                                                                 (5, 4, 32, False, True)])
def test_gru_cell_head_forward_backward(ops, b, steps, hidden, learn_origin, with_ts):
  """tfpp_gru_cell_head / _bwd == nn.GRUCell unrolled (model.py:886-913) + target-speed MLP, values and BPTT gradients."""
  g = torch.Generator().manual_seed(7)
  cell = torch.nn.GRUCell(4, hidden).double()
  out = torch.nn.Linear(hidden, 2).double()
  ts0, ts1 = torch.nn.Linear(hidden, hidden).double(), torch.nn.Linear(hidden, 4).double()
  js = hidden + (2 if learn_origin else 0)
  joined = torch.rand(b, js, generator=g, dtype=torch.float64).requires_grad_(True)
  tp = torch.randn(b, 2, generator=g, dtype=torch.float64) * 5
  z = joined[:, :hidden]
  x = joined[:, hidden:hidden + 2] if learn_origin else torch.zeros(b, 2, dtype=torch.float64)
  wps = []
  for _ in range(steps):
    z = cell(torch.cat([x, tp], 1), z)
    x = out(z) + x
    wps.append(x)
  want_wp = torch.stack(wps, 1)
  want_ts = ts1(F.relu(ts0(joined[:, :hidden]))) if with_ts else None
  dwp = torch.randn(b, steps, 2, generator=g, dtype=torch.float64)
  dts = torch.randn(b, 4, generator=g, dtype=torch.float64) if with_ts else None
  ((want_wp * dwp).sum() + ((want_ts * dts).sum() if with_ts else 0.0)).backward()
  f = lambda t: t.detach().float().cuda().contiguous()
  W = [f(cell.weight_ih), f(cell.weight_hh), f(cell.bias_ih), f(cell.bias_hh), f(out.weight), f(out.bias)]
  T = [f(ts0.weight), f(ts0.bias), f(ts1.weight), f(ts1.bias)] if with_ts else [None] * 4
  jd = f(joined)
  wp, ts, h_all = ops.gru_cell_head(jd, f(tp), *W, *T, steps=steps, hidden=hidden, learn_origin=learn_origin, want_h=True)
  assert rel(wp, want_wp) < 1e-5 and (not with_ts or rel(ts, want_ts) < 1e-5)
  grads = [torch.zeros_like(t) for t in W] + ([torch.zeros_like(t) for t in T] if with_ts else [None] * 4)
  dj = torch.zeros_like(jd)
  ops.gru_cell_head_bwd(jd, f(tp), *W, *T[:3], wp, h_all, f(dwp), f(dts) if with_ts else None, dj, grads, steps=steps,
                        hidden=hidden, learn_origin=learn_origin)
  assert rel(dj, joined.grad) < 1e-4
  for got, p in zip(grads[:6], (cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh, out.weight, out.bias)):
    assert rel(got, p.grad) < 1e-4
  if with_ts:
    for got, p in zip(grads[6:], (ts0.weight, ts0.bias, ts1.weight, ts1.bias)):
      assert rel(got, p.grad) < 1e-4
  # accumulation semantics: a second call doubles everything
  ops.gru_cell_head_bwd(jd, f(tp), *W, *T[:3], wp, h_all, f(dwp), f(dts) if with_ts else None, dj, grads, steps=steps,
                        hidden=hidden, learn_origin=learn_origin)
  assert rel(dj, 2 * joined.grad) < 1e-4 and rel(grads[1], 2 * cell.weight_hh.grad) < 1e-4


def _torch_losses_wp(model, out, lab):
  from tests.test_boundary_gpu import _torch_losses
  loss = _torch_losses(model, out, lab)
  loss['loss_wp'] = torch.mean(torch.abs(out[0] - lab['waypoint']))
  return loss


def _data():
  from carla_garage_b200 import synth
  inp = {k: v.cuda() for k, v in synth.make_inputs(2, seed=11).items()}
  lab = {k: v.cuda().contiguous() for k, v in synth.make_labels(2, seed=13).items()}
  lab['waypoint'] = synth.make_waypoint_labels(2, 8, seed=13).cuda()
  return inp, lab


@pytest.mark.noisy
def test_mlp_join_train_step_fp32_vs_reference_golden(ops):
  """fp32 parity mode through the autograd boundary: outputs, the eleven losses and the gradients of every planner
  parameter (join MLP, both GRUCell heads, target-speed MLP, extra-sensor encoder, lidar_to_img_features_end) within
  1e-3 of the unmodified reference."""
  from carla_garage_b200 import synth
  from carla_garage_b200.nn import LidarCenterNet
  g = np.load(os.path.join(GOLDEN, 'mlp_join_b2.npz'))
  with ops.precision('fp32'):
    m = LidarCenterNet(_cfg())
    m.load_state_dict(synth.mlp_join_state(GOLDEN), strict=True)
    m = m.cuda().train()
    inp, lab = _data()
    out = m(**inp)
    assert out[0].shape == (2, 8, 2) and out[2].shape == (2, 10, 2) and out[1].shape == (2, 4)
    errs = {'pred_wp': rel(out[0], g['pred_wp']), 'pred_target_speed': rel(out[1], g['pred_target_speed']),
            'pred_checkpoint': rel(out[2], g['pred_checkpoint'])}
    losses = _torch_losses_wp(m, out, lab)
    assert len(losses) == 11
    for k, v in losses.items():
      errs[k] = abs(float(v) - float(g[k])) / max(abs(float(g[k])), 1e-6)
    (sum(losses.values()) / len(losses)).backward()
    torch.cuda.synchronize()
  params = dict(m.named_parameters())
  n = 0
  for key in g.files:
    if key.startswith('grad_'):
      name = key[5:]
      errs['grad ' + name] = rel(params[name].grad.flatten()[:512], g[key])
      n += 1
  print('\n' + '\n'.join(f'  fp32 mlp-join {k}: {v:.2e}' for k, v in errs.items()))
  assert n >= 20
  for k, v in errs.items():
    assert v < 1e-3, (k, v)


@pytest.mark.noisy
def test_mlp_join_bf16_trainer_step_and_graph(ops):
  """Production precision: the fused Trainer step (losses incl. loss_wp) tracks the reference within the bf16 floor of
  this network, and the captured graph replays it."""
  from carla_garage_b200 import synth
  from carla_garage_b200.nn import LidarCenterNet
  from carla_garage_b200.training import Trainer
  g = np.load(os.path.join(GOLDEN, 'mlp_join_b2.npz'))
  m = LidarCenterNet(_cfg())
  m.load_state_dict(synth.mlp_join_state(GOLDEN), strict=True)
  tr = Trainer(m.cuda().train())
  assert 'loss_wp' in tr.keys and len(tr.keys) == 11
  inp, lab = _data()
  out, losses = tr.forward_backward(inp, lab)
  torch.cuda.synchronize()
  for k in ('loss_wp', 'loss_checkpoint', 'loss_target_speed', 'loss_semantic'):
    assert abs(float(losses[k]) - float(g[k])) <= 0.15 * max(abs(float(g[k])), 0.05), (k, float(losses[k]), float(g[k]))
  assert rel(out[0], g['pred_wp']) < 0.15 and rel(out[2], g['pred_checkpoint']) < 0.15
  p = dict(m.named_parameters())
  for name in ('join.0.weight', 'checkpoint_decoder.wp_decoder.weight_hh', 'wp_decoder.output.weight',
               'backbone.lidar_to_img_features_end.weight'):
    a, b = p[name].grad.flatten()[:512].double().cpu(), torch.from_numpy(g['grad_' + name]).double()
    assert float((a * b).sum() / (a.norm() * b.norm() + 1e-30)) > 0.9, name
  tr.capture(inp, lab)
  l0 = float(tr.replay()[1].sum())
  l1 = float(tr.replay()[1].sum())
  assert tr.graph_opt is None and l0 == l0 and l1 < l0   # one graph; same batch twice: the loss goes down
