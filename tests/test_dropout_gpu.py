"""Training-mode dropout (transfuser.py:325,374,379,395; nn.TransformerDecoderLayer / nn.MultiheadAttention dropout 0.1,
model.py:137-140) on the B200 kernels.  torch's dropout masks come from its own generator and cannot be reproduced, so
parity is established the other way round: the kernels' counter-based Philox4x32-10 stream is restated in numpy
(oracle/philox.py, pinned to the Random123 known-answer vectors) and the SAME masks are applied at the reference's
dropout sites inside the fp32 oracle."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
SEED = 0x1234ABCD5678


@pytest.fixture(scope='module')
def ops():
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  from carla_garage_b200 import ops as o
  return o


def rel(a, b):
  a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
  return float((a - b).norm() / (b.norm() + 1e-30))


def rnd(*shape, seed=0, scale=1.0):
  g = torch.Generator().manual_seed(seed)
  return (torch.randn(*shape, generator=g) * scale).cuda()


def bf(x):
  return x.to(torch.bfloat16)


def rng_tensor(step=1):
  return torch.tensor([SEED, step], dtype=torch.int64, device='cuda')


def mult(shape, p, step, site):
  from oracle import philox
  return philox.multiplier(shape, p, SEED, step, site).cuda()


def test_philox_stream_is_the_oracle_stream(ops):
  for n, p, step, site in ((8, 0.1, 1, 0), (4096 + 8, 0.1, 7, 3), (1 << 16, 0.5, 2**31 + 5, 41), (1024, 0.0, 1, 1)):
    x = torch.ones(n, device='cuda')
    ops.dropout_(x, (rng_tensor(step), p, site))
    want = mult((n,), p, step, site) if p > 0 else torch.ones(n, device='cuda')
    assert torch.equal(x, want), (n, p, step, site)
  xb = bf(rnd(2048, seed=1))
  want = (xb.float() * mult((2048,), 0.1, 3, 9)).to(torch.bfloat16)
  ops.dropout_(xb, (rng_tensor(3), 0.1, 9))
  assert torch.equal(xb, want)
  m = mult((1 << 16,), 0.1, 5, 0)
  assert abs(float((m == 0).float().mean()) - 0.1) < 0.01
  assert not torch.equal(mult((4096,), 0.1, 5, 0), mult((4096,), 0.1, 6, 0))   # another step, another mask
  assert not torch.equal(mult((4096,), 0.1, 5, 0), mult((4096,), 0.1, 5, 1))   # another site, another mask


@pytest.mark.parametrize('rows,k,n', [(640, 216, 216), (352, 2048, 256), (320, 1512, 1512), (100, 72, 72)])
def test_gemm_epilogue_dropout(ops, rows, k, n):
  x, w, b = bf(rnd(rows, k, seed=2)), bf(rnd(n, k, seed=3, scale=k**-0.5)), rnd(n, seed=4)
  res = rnd(rows, n, seed=5)
  pre = x.float() @ w.float().t() + b
  d = (rng_tensor(4), 0.1, 17)
  got = ops.linear(x, w, bias=b, res=res, out_f32=True, drop=d)       # x + drop(proj(y)): transfuser.py:379
  want = pre * mult((rows, n), 0.1, 4, 17) + res
  assert rel(got, want) < 2e-3
  got = ops.linear(x, w, bias=b, act=ops.ACT_RELU, drop=d)            # dropout(activation(linear1(x))): decoder FFN
  want = F.relu(pre) * mult((rows, n), 0.1, 4, 17)
  assert rel(got.float(), want) < 6e-3
  assert float(((got.float() == 0) != (want == 0)).float().mean()) < 1e-3
  # adjoint: dz = mask * dy (+ bias gradient)
  dy = rnd(rows, n, seed=6)
  db = torch.zeros(n, device='cuda')
  dz = ops.act_bwd(dy, None, ops.ACT_NONE, 1, rows, n, layout=2, dbias=db, drop=d)
  want = dy * mult((rows, n), 0.1, 4, 17)
  assert rel(dz.float(), want) < 4e-3 and rel(db, want.sum(0)) < 1e-4
  dz = ops.act_bwd(bf(dy), None, ops.ACT_NONE, 1, rows, n, layout=0, drop=d)
  assert rel(dz.float(), bf(dy).float() * mult((rows, n), 0.1, 4, 17)) < 4e-3


@pytest.mark.parametrize('c,b,t', [(72, 2, 320), (216, 2, 320), (576, 1, 320), (1512, 1, 320), (72, 3, 64)])
def test_fusion_attention_dropout_forward_backward(ops, c, b, t):
  heads, p, step, site = 4, 0.1, 3, 5
  hd = c // heads
  qkv = bf(rnd(b, t, 3 * c, seed=7))
  d = (rng_tensor(step), p, site)
  out = ops.fusion_attn(qkv, b, t, c, heads, drop=d).view(b, t, c)
  leaf = qkv.double().clone().requires_grad_(True)
  q, k, v = [u.view(b, t, heads, hd).transpose(1, 2) for u in leaf.split(c, dim=2)]
  att = F.softmax(q @ k.transpose(-2, -1) / math.sqrt(hd), dim=-1) * mult((b, heads, t, t), p, step, site).double()
  ref = (att @ v).transpose(1, 2).reshape(b, t, c)
  assert rel(out.float(), ref) < 1e-2
  dout = bf(rnd(b * t, c, seed=8))
  ref.backward(dout.double().view(b, t, c))
  dqkv = ops.fusion_attn_bwd(qkv, dout, b, t, c, heads, drop=d)
  assert rel(dqkv.float(), leaf.grad) < 2e-2
  # p = 0 / no dropout is the plain kernel
  assert torch.equal(ops.fusion_attn(qkv, b, t, c, heads), ops.fusion_attn(qkv, b, t, c, heads, drop=None))


@pytest.mark.parametrize('tq,tk,cross', [(11, 11, False), (11, 65, True)])
def test_small_mha_dropout_forward_backward(ops, tq, tk, cross):
  b, heads, hd, p, step, site = 3, 8, 32, 0.1, 9, 2
  dm = heads * hd
  d = (rng_tensor(step), p, site)
  if cross:
    q, kv = bf(rnd(b, tq, dm, seed=9)), bf(rnd(b, tk, 2 * dm, seed=10))
    qs, ks, vs = (tq * dm, dm), (tk * 2 * dm, 2 * dm), (tk * 2 * dm, 2 * dm)
    out = ops.small_mha(q, kv, kv, b, heads, tq, tk, hd, qs, ks, vs, v_off=dm, drop=d)
    qf, kf, vf = q.double(), kv.double()[..., :dm], kv.double()[..., dm:]
  else:
    qkv = bf(rnd(b, tq, 3 * dm, seed=11))
    s3 = (tq * 3 * dm, 3 * dm)
    out = ops.small_mha(qkv, qkv, qkv, b, heads, tq, tq, hd, s3, s3, s3, k_off=dm, v_off=2 * dm, drop=d)
    qf, kf, vf = qkv.double()[..., :dm], qkv.double()[..., dm:2 * dm], qkv.double()[..., 2 * dm:]
  qf, kf, vf = (u.clone().requires_grad_(True) for u in (qf, kf, vf))
  split = lambda u, n: u.reshape(b, n, heads, hd).transpose(1, 2)
  att = F.softmax(split(qf, tq) @ split(kf, tk).transpose(-2, -1) / math.sqrt(hd), -1)
  att = att * mult((b, heads, tq, tk), p, step, site).double()
  ref = (att @ split(vf, tk)).transpose(1, 2).reshape(b * tq, dm)
  assert rel(out.float(), ref) < 5e-3
  dout = bf(rnd(b * tq, dm, seed=12))
  ref.backward(dout.double())
  if cross:
    dq = torch.empty(b * tq, dm, dtype=torch.bfloat16, device='cuda')
    dkv = torch.empty_like(kv)
    ops.small_mha_bwd(q, kv, kv, dout, dq, dkv, dkv, b, heads, tq, tk, hd, qs, ks, vs, (tq * dm, dm), ks, vs,
                      offs=(0, 0, dm, 0, 0, dm), drop=d)
    assert rel(dq.float(), qf.grad.reshape(b * tq, dm)) < 1e-2
    assert rel(dkv.float()[..., :dm], kf.grad) < 1e-2 and rel(dkv.float()[..., dm:], vf.grad) < 1e-2
  else:
    dqkv = torch.empty_like(qkv)
    ops.small_mha_bwd(qkv, qkv, qkv, dout, dqkv, dqkv, dqkv, b, heads, tq, tq, hd, s3, s3, s3, s3, s3, s3,
                      offs=(0, dm, 2 * dm, 0, dm, 2 * dm), drop=d)
    assert rel(dqkv.float(), torch.cat([qf.grad, kf.grad, vf.grad], dim=-1)) < 1e-2


@pytest.fixture(scope='module')
def trainer(oracle_state):
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  from carla_garage_b200.config import GlobalConfig
  from carla_garage_b200.nn import LidarCenterNet
  from carla_garage_b200.training import Trainer
  m = LidarCenterNet(GlobalConfig())
  m.load_state_dict(oracle_state, strict=True)
  tr = Trainer(m.cuda().train())
  tr.eng.dropout_enabled = True
  return tr


def _grad_state(oracle_state, gemm_prefixes):
  gemm = lambda k, v: v.dim() >= 2 and k.startswith(gemm_prefixes)
  return {k: ((v.to(torch.bfloat16).float() if gemm(k, v) else v.clone()).requires_grad_(True)
              if v.is_floating_point() and 'running' not in k else v) for k, v in oracle_state.items()}


@pytest.mark.noisy
def test_fusion_block_with_dropout_vs_oracle(ops, trainer, oracle_state):
  """fuse_features + GPT (transfuser.py:222-257,301-339) in training mode with all four dropout kinds active, forward and
  backward, against the fp32 oracle given the same masks."""
  from carla_garage_b200.training import Backward
  from oracle import philox, tfpp_oracle as orc
  net, eng, st = trainer.model, trainer.eng, trainer.st
  net.load_state_dict(oracle_state, strict=True)
  sd = _grad_state(oracle_state, ('backbone.transformers.1.', 'backbone.lidar_channel_to_img.1.',
                                  'backbone.img_channel_to_lidar.1.'))
  g = torch.Generator().manual_seed(21)
  xi = bf(torch.randn(2, 216, 32, 128, generator=g)).float().requires_grad_(True)
  xl = bf(torch.randn(2, 216, 32, 32, generator=g)).float().requires_grad_(True)
  eng.seed_dropout(SEED, torch.device('cuda'), step=4)
  eng.begin_dropout_step(torch.device('cuda'))   # -> step 5, sites from 0
  stream = philox.DropoutStream(SEED, 5, 0)
  wi, wl = orc.fuse_features(sd, 'backbone', xi, xl, 1, orc.DEFAULT_CFG, dropout=stream)
  assert stream.site == 1 + 2 * 3   # embd + 2 blocks x (attn, proj, mlp)
  di, dl = torch.randn(wi.shape, generator=g), torch.randn(wl.shape, generator=g)
  (wi * di).sum().add((wl * dl).sum()).backward()
  to_dev = lambda t: ops.nchw_to_nhwc(t.detach().cuda().contiguous())
  st.zero_grad()
  eng.tape = []
  try:
    xid, xld = to_dev(xi), to_dev(xl)
    yi, yl = eng.fuse(xid, xld, 1, True)
    tape = eng.tape
  finally:
    eng.tape = None
  assert eng._site == stream.site  # pylint: disable=protected-access
  assert rel(ops.nhwc_to_nchw(yi), wi) < 1e-2 and rel(ops.nhwc_to_nchw(yl), wl) < 1e-2
  bw = Backward(eng, st)
  bw.G[id(yi)], bw.G[id(yl)] = to_dev(di), to_dev(dl)
  bw.run(tape, {})
  torch.cuda.synchronize()
  assert rel(ops.nhwc_to_nchw(bw.G[id(xid)]), xi.grad) < 3e-2 and rel(ops.nhwc_to_nchw(bw.G[id(xld)]), xl.grad) < 3e-2
  params = dict(net.named_parameters())
  for n in sd:
    if not n.startswith(('backbone.transformers.1.', 'backbone.lidar_channel_to_img.1.', 'backbone.img_channel_to_lidar.1.')):
      continue
    if n.endswith('attn.key.bias') or sd[n].grad is None:
      continue
    assert rel(params[n].grad, sd[n].grad) < 5e-2, n
  # without the masks the oracle is far away: the comparison above really exercises the dropout path
  wi0, _ = orc.fuse_features(sd, 'backbone', xi, xl, 1, orc.DEFAULT_CFG)
  assert rel(ops.nhwc_to_nchw(yi), wi0) > 3e-2


@pytest.mark.noisy
def test_planner_with_dropout_vs_oracle(ops, trainer, oracle_state):
  """6-layer post-norm decoder with the six dropouts of nn.TransformerDecoderLayer per layer (model.py:137-140)."""
  from carla_garage_b200.training import Backward
  from oracle import philox, tfpp_oracle as orc
  net, eng, st = trainer.model, trainer.eng, trainer.st
  net.load_state_dict(oracle_state, strict=True)
  sd = _grad_state(oracle_state, ('join.', 'change_channel'))
  b = 4
  g = torch.Generator().manual_seed(22)
  fused = bf(torch.randn(b, 1512, 8, 8, generator=g)).float().requires_grad_(True)
  tp, vel = torch.randn(b, 2, generator=g) * 10, torch.rand(b, 1, generator=g) * 8
  cmd = F.one_hot(torch.randint(0, 6, (b,), generator=g), 6).float()
  eng.seed_dropout(SEED, torch.device('cuda'), step=10)
  eng.begin_dropout_step(torch.device('cuda'))
  stream = philox.DropoutStream(SEED, 11, 0)
  want_cp, want_ts = orc.planner(sd, fused, tp, vel, cmd, training=True, dropout=stream)
  assert stream.site == 36
  dcp, dts = torch.randn(want_cp.shape, generator=g), torch.randn(want_ts.shape, generator=g)
  (want_cp * dcp).sum().add((want_ts * dts).sum()).backward()
  st.zero_grad()
  eng.tape = []
  eng.new_arena(torch.device('cuda'))
  try:
    xd = ops.nchw_to_nhwc(fused.detach().cuda().contiguous())
    cp, ts, _ = eng.planner(xd, tp.cuda(), vel.cuda(), cmd.cuda(), True)
    tape = eng.tape
  finally:
    eng.tape = None
  assert eng._site == 36  # pylint: disable=protected-access
  assert rel(cp, want_cp) < 1.5e-2 and rel(ts, want_ts) < 1.5e-2
  bw = Backward(eng, st)
  bw.run(tape, {'planner': (dcp.cuda(), dts.cuda())})
  torch.cuda.synchronize()
  assert rel(ops.nhwc_to_nchw(bw.G[id(xd)]), fused.grad) < 6e-2
  params = dict(net.named_parameters())
  names = [n for n in sd if n.startswith(('join.', 'change_channel', 'checkpoint_', 'target_speed_network',
                                          'extra_sensor_')) and sd[n].is_floating_point() and sd[n].grad is not None]
  for n in names:
    assert rel(params[n].grad, sd[n].grad) < 8e-2, n
  net.load_state_dict(oracle_state, strict=True)


@pytest.mark.noisy
def test_train_step_with_dropout_and_graph_replay(trainer, oracle_state):
  """The whole step with dropout on: every replay of the captured graph draws a new mask (the step counter lives on the
  device), eval stays deterministic, and the loss still goes down on a repeated batch."""
  from carla_garage_b200 import synth
  net, eng = trainer.model, trainer.eng
  net.load_state_dict(oracle_state, strict=True)
  net.train()
  inp = {k: v.cuda() for k, v in synth.make_inputs(2, seed=11).items()}
  lab = {k: v.cuda().contiguous() for k, v in synth.make_labels(2, seed=13).items()}
  eng.seed_dropout(SEED, torch.device('cuda'))
  _, l0 = trainer.step(inp, lab)
  assert eng._site == 4 * 7 + 36  # 4 fusion scales x (embd + 2 x 3) + 6 decoder layers x 6  pylint: disable=protected-access
  assert int(eng.rng[1]) == 1
  trainer.capture(inp, lab)
  s0 = int(eng.rng[1])
  vals = []
  for _ in range(3):
    _, gl = trainer.replay()
    vals.append(gl.clone())
  torch.cuda.synchronize()
  assert int(eng.rng[1]) == s0 + 3
  assert all(bool(torch.isfinite(v).all()) for v in vals)
  assert float(vals[2].sum()) < float(sum(float(x) for x in l0.values()))
  # eval mode draws no masks: two eval forwards agree up to the run-to-run noise of the fp32 atomics (a training
  # forward with a 10 % dropout would move the outputs by far more than that)
  net.eval()
  with torch.no_grad():
    a = net(**inp)
    b2 = net(**inp)
  assert rel(a[2], b2[2]) < 2e-2
  net.train()
