"""Run the mlp-join Trainer step several times in one process and report the first backward tensor that differs between runs."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('TFPP_DROPOUT', '0')
from carla_garage_b200 import ops, synth, training
from carla_garage_b200.config import GlobalConfig
from carla_garage_b200.nn import LidarCenterNet
from carla_garage_b200.training import Trainer
G = os.path.join(ROOT, 'tests', 'golden')
log = {}
orig_act_bwd, orig_linear, orig_wgrad = ops.act_bwd, ops.linear, ops.conv_wgrad
state = {'on': False, 'n': 0}
def wrap(name, fn):
  def f(*a, **k):
    out = fn(*a, **k)
    if state['on']:
      torch.cuda.synchronize()
      ins = [t for t in list(a) + list(k.values()) if torch.is_tensor(t)]
      log.setdefault(state['run'], []).append((name, state['n'], [float(t.float().abs().sum()) for t in ins[:3]],
                                                float(out.float().abs().sum()) if torch.is_tensor(out) else None))
      state['n'] += 1
    return out
  return f
ops.act_bwd, ops.linear, ops.conv_wgrad = wrap('act_bwd', orig_act_bwd), wrap('linear', orig_linear), wrap('wgrad', orig_wgrad)
orig_mlp = training.Backward.mlp_join
orig_gf = training.Backward.global_fuse
def mlp(self, r):
  state['on'] = True
  try:
    return orig_mlp(self, r)
  finally:
    state['on'] = False
def gf(self, r):
  state['on'] = True
  try:
    return orig_gf(self, r)
  finally:
    state['on'] = False
training.Backward.mlp_join, training.Backward.global_fuse = mlp, gf
cfg = GlobalConfig(); cfg.transformer_decoder_join = False; cfg.use_wp_gru = True
res = []
for run in range(6):
  state['run'], state['n'] = run, 0
  m = LidarCenterNet(cfg); m.load_state_dict(synth.mlp_join_state(G), strict=True)
  tr = Trainer(m.cuda().train())
  inp = {k: v.cuda() for k, v in synth.make_inputs(2, seed=11).items()}
  lab = {k: v.cuda().contiguous() for k, v in synth.make_labels(2, seed=13).items()}
  lab['waypoint'] = synth.make_waypoint_labels(2, 8, seed=13).cuda()
  tr.forward_backward(inp, lab)
  torch.cuda.synchronize()
  p = dict(m.named_parameters())
  res.append({n: float(p[n].grad.float().abs().sum()) for n in ('join.0.weight', 'join.2.weight', 'join.4.weight',
                                                                 'backbone.lidar_to_img_features_end.weight',
                                                                 'checkpoint_decoder.wp_decoder.weight_hh', 'wp_decoder.output.weight')})
  print('run', run, {k: round(v, 4) for k, v in res[-1].items()})
  del tr, m
for run in range(6):
  for a, b in zip(log[0], log[run]):
    if a[0] != b[0] or any(abs(x - y) > 1e-2 * (abs(x) + 1e-6) for x, y in zip(a[2], b[2])) or (a[3] is not None and abs(a[3] - b[3]) > 1e-2 * (abs(a[3]) + 1e-6)):
      print('run', run, 'first divergence vs run 0:', a, b)
      break
  else:
    print('run', run, 'matches run 0 on', len(log[run]), 'ops')
