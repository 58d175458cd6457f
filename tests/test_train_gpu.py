"""GPU parity of the training step: fused losses and parameter gradients against golden values produced by the
unmodified reference's loss.backward() (tests/golden/make_golden.py, train_b2.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def rel(a, b):
  a, b = torch.as_tensor(a).double().cpu().flatten(), torch.as_tensor(b).double().cpu().flatten()
  return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope='module')
def trainer(oracle_state):
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  from carla_garage_b200.config import GlobalConfig
  from carla_garage_b200.nn import LidarCenterNet
  from carla_garage_b200.training import Trainer
  m = LidarCenterNet(GlobalConfig())
  m.load_state_dict(oracle_state, strict=True)
  m = m.cuda().train()
  return Trainer(m)


def test_losses_and_gradients_vs_reference(trainer):
  from carla_garage_b200 import synth
  g = np.load(os.path.join(GOLDEN, 'train_b2.npz'))
  ge = np.load(os.path.join(GOLDEN, 'forward_eval_b2.npz'))
  inp = {k: v.cuda() for k, v in synth.make_inputs(2, seed=11).items()}
  lab = {k: v.cuda().contiguous() for k, v in synth.make_labels(2, seed=13).items()}
  out, losses = trainer.forward_backward(inp, lab)
  torch.cuda.synchronize()
  floor = max(0.02, 2.5 * float(ge['bf16floor_pred_semantic']))
  report = {}
  for k, v in losses.items():
    report[k] = (float(v), float(g[k]))
  print('\n'.join(f'  {k}: got {a:.5f} want {b:.5f}' for k, (a, b) in report.items()))
  for k, (a, b) in report.items():
    assert abs(a - b) <= floor * max(abs(b), 0.05), (k, a, b)
  total = sum(float(v) for v in losses.values()) / 10
  assert abs(total - float(g['total'])) <= floor * float(g['total'])
  params = dict(trainer.model.named_parameters())
  errs = {}
  for key in g.files:
    if not key.startswith('gradnorm_'):
      continue
    name = key[len('gradnorm_'):]
    gr = params[name].grad
    errs[name] = (float(gr.norm()) / float(g[key]), rel(gr.flatten()[:256], g['grad_' + name]))
  print('\n'.join(f'  {k}: norm ratio {a:.3f}  first-256 rel err {b:.3f}' for k, (a, b) in errs.items()))
  # heads / decoder (short bf16 paths) must be tight; deep backbone gradients are bounded by the bf16 noise floor
  for k, (ratio, e) in errs.items():
    assert 0.5 < ratio < 2.0, (k, ratio)
  tight = [k for k in errs if k.startswith(('head.', 'semantic_decoder.deconv3', 'target_speed_network'))]
  for k in tight:
    assert errs[k][1] < 0.1, (k, errs[k])


def test_adamw_step_matches_torch(trainer):
  from carla_garage_b200 import synth
  st = trainer.st
  st.zero_grad()
  gen = torch.Generator(device='cuda').manual_seed(0)
  p = st.params[5]
  before = p.detach().clone()
  st.grad.normal_(generator=gen)
  ref_p = before.clone().requires_grad_(True)
  opt = torch.optim.AdamW([ref_p], lr=3e-4, amsgrad=True)
  for step in range(3):
    ref_p.grad = st.g(p).clone()
    opt.step()
    st.adamw_step(3e-4)
  assert rel(p.detach(), ref_p.detach()) < 1e-6
  del synth
