"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/team_code) in the build
container.  The reference cannot travel to the GPU box, so its outputs are committed as small fixtures.

  python tests/golden/make_golden.py

Inputs/weights come from carla_garage_b200.synth (pure functions of a seed), so the same tensors can be rebuilt on
the GPU box and pushed through the CUDA path and the oracle.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from carla_garage_b200 import compat, synth  # noqa: E402
from oracle.regnety import timm_factory  # noqa: E402
from tests.golden.sampling import sample  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def sub(t):
  return sample(t.detach()).numpy()


def main():
  compat.install(timm_factory)
  from config import GlobalConfig  # pylint: disable=import-outside-toplevel
  from model import LidarCenterNet  # pylint: disable=import-outside-toplevel
  torch.manual_seed(0)
  torch.set_num_threads(os.cpu_count())
  cfg = GlobalConfig()
  net = LidarCenterNet(cfg)
  ref_sd = net.state_dict()
  shapes = {k: list(v.shape) for k, v in ref_sd.items()}
  dtypes = {k: str(v.dtype) for k, v in ref_sd.items()}
  json.dump(dict(shapes=shapes, dtypes=dtypes), open(os.path.join(OUT, 'state_dict_keys.json'), 'w'), indent=0)
  fixed = {k: ref_sd[k] for k in ('valid_bev_pixels', 'valid_bev_pixels_inv', 'loss_speed.weight',
                                  'loss_semantic.weight', 'loss_bev_semantic.weight')}
  np.savez_compressed(os.path.join(OUT, 'valid_bev_pixels.npz'),
                      valid=ref_sd['valid_bev_pixels'].numpy().astype(np.uint8))

  # ---- K1: lidar_to_histogram_features (data.py:873-906) ----
  pts = synth.make_point_clouds(2, seed=7).numpy()
  edge = np.array([[32.0, 0.0, 1.0], [32.0001, 0.0, 1.0], [-32.0, -32.0, 1.0], [0.0, 32.0, 1.0], [-32.0001, 0, 1],
                   [1.0, 1.0, 0.2], [1.0, 1.0, 0.20001], [2.0, 2.0, 100.0], [2.0, 2.0, 99.9], [-0.0, -0.0, 3.0],
                   [31.99999, 31.99999, 3.0], [-1e-7, 1e-7, 3.0]] + [[5.1, 5.1, 1.0]] * 9, dtype=np.float32)
  k1 = {}
  for name, cloud in (('cloud0', pts[0]), ('cloud1', pts[1]), ('edge', edge), ('empty', np.zeros((0, 3), np.float32))):
    for gp in (False, True):
      out = net.data.lidar_to_histogram_features(cloud, use_ground_plane=gp)
      k1[f'{name}_gp{int(gp)}'] = np.round(out * 5).astype(np.uint8)
  np.savez_compressed(os.path.join(OUT, 'pillar_scatter.npz'), edge_points=edge, **k1)

  # ---- full forward, eval mode ----
  sd = synth.make_state_dict(shapes, seed=0, fixed=fixed)
  net.load_state_dict(sd, strict=True)
  B = 2
  inp = synth.make_inputs(B, seed=11)
  # Calibrate the BatchNorm running statistics like a trained checkpoint's would be (running stats == statistics of
  # the data): one training-mode pass with momentum 1.  Purely random running stats make the eval-mode residual
  # stream grow ~1.4x per block (7.8e5 after stage 3), an ill-conditioned network no bf16 path can track.
  net.train()
  bns = [m for m in net.modules() if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d))]
  for m in bns:
    m.momentum = 1.0
  for m in net.modules():
    if isinstance(m, torch.nn.Dropout):
      m.p = 0.0
    if isinstance(m, torch.nn.MultiheadAttention):
      m.dropout = 0.0
  with torch.no_grad():
    net(**inp)  # calibrate on the evaluation batch itself: eval-mode statistics == batch statistics
  for m in bns:
    m.momentum = 0.1
  calib = {k: v.clone() for k, v in net.state_dict().items() if k.endswith('running_mean') or k.endswith('running_var')}
  np.savez_compressed(os.path.join(OUT, 'bn_calib.npz'), **{k: v.numpy() for k, v in calib.items()})
  sd.update(calib)
  net.load_state_dict(sd, strict=True)
  taps = {}
  hooks = []
  bb = net.backbone

  def hook(name):
    return lambda m, i, o: taps.__setitem__(name, o.detach())

  hooks.append(bb.image_encoder['stem'].register_forward_hook(hook('img_stem')))
  hooks.append(bb.lidar_encoder['stem'].register_forward_hook(hook('lid_stem')))
  for i in range(4):
    hooks.append(bb.image_encoder[f's{i + 1}'].register_forward_hook(hook(f'img_s{i + 1}_pre')))
    hooks.append(bb.lidar_encoder[f's{i + 1}'].register_forward_hook(hook(f'lid_s{i + 1}_pre')))
  hooks.append(net.join.register_forward_hook(hook('joined')))
  orig_backbone_forward = bb.forward

  def wrapped(image, lidar):
    feats = orig_backbone_forward(image, lidar)
    taps['bev_feature_grid'], taps['fused_features'], taps['image_feature_grid'] = [f.detach() for f in feats]
    return feats

  bb.forward = wrapped
  net.eval()
  with torch.no_grad():
    out = net(**inp)
  g = {}
  g['pred_target_speed'] = out[1].numpy()
  g['pred_checkpoint'] = out[2].numpy()
  g['pred_semantic'] = sub(out[3])
  g['pred_bev_semantic'] = sub(out[4])
  g['pred_depth'] = sub(out[5])
  for n, t in zip(('heatmap', 'wh', 'offset', 'yaw_class', 'yaw_res'), out[6][:5]):
    g['bb_' + n] = sub(t)
    g['norm_bb_' + n] = np.array(float(t.norm()))
  for k, t in taps.items():
    g['tap_' + k] = sub(t)
    g['norm_' + k] = np.array(float(t.norm()))
  for n, t in (('pred_semantic', out[3]), ('pred_bev_semantic', out[4]), ('pred_depth', out[5])):
    g['norm_' + n] = np.array(float(t.norm()))
  boxes = net.head.get_bboxes(*out[6])
  g['boxes'] = boxes.numpy()
  # noise floor of bf16 storage on this state (oracle/bf16_emulation.py); the oracle == reference to 1e-6
  from oracle import bf16_emulation  # pylint: disable=import-outside-toplevel
  etaps = {}
  eout = bf16_emulation.forward(sd, inp, taps=etaps)

  def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))

  for k, t in etaps.items():
    if k in taps:
      g['bf16floor_' + k] = np.array(rel(t, taps[k]))
  for n, i in (('pred_target_speed', 1), ('pred_checkpoint', 2), ('pred_semantic', 3), ('pred_bev_semantic', 4),
               ('pred_depth', 5)):
    g['bf16floor_' + n] = np.array(rel(eout[i], out[i]))
  for n, a, b in zip(('heatmap', 'wh', 'offset', 'yaw_class', 'yaw_res'), eout[6][:5], out[6][:5]):
    g['bf16floor_bb_' + n] = np.array(rel(a, b))
  print({k: float(v) for k, v in g.items() if k.startswith('bf16floor_')})
  np.savez_compressed(os.path.join(OUT, 'forward_eval_b2.npz'), **g)

  # ---- train mode (batch-stat BN), dropout disabled, losses + a few gradients ----
  net.train()
  for m in net.modules():
    if isinstance(m, torch.nn.Dropout):
      m.p = 0.0
    if isinstance(m, torch.nn.MultiheadAttention):
      m.dropout = 0.0
  lab = synth.make_labels(B, seed=13)
  net.load_state_dict(sd, strict=True)
  out = net(**inp)
  losses = net.compute_loss(pred_wp=out[0], pred_target_speed=out[1], pred_checkpoint=out[2], pred_semantic=out[3],
                            pred_bev_semantic=out[4], pred_depth=out[5], pred_bounding_box=out[6], pred_wp_1=out[8],
                            selected_path=out[9], waypoint_label=None, target_speed_label=lab['target_speed'],
                            checkpoint_label=lab['checkpoint'], semantic_label=lab['semantic'],
                            bev_semantic_label=lab['bev_semantic'], depth_label=lab['depth'],
                            center_heatmap_label=lab['center_heatmap'], wh_label=lab['wh'],
                            yaw_class_label=lab['yaw_class'], yaw_res_label=lab['yaw_res'],
                            offset_label=lab['offset'], velocity_label=None, brake_target_label=None,
                            pixel_weight_label=lab['pixel_weight'], avg_factor_label=lab['avg_factor'])
  total = sum(losses.values()) / len(losses)
  total.backward()
  t = {k: np.array(float(v)) for k, v in losses.items()}
  t['total'] = np.array(float(total))
  t['pred_target_speed'] = out[1].detach().numpy()
  t['pred_checkpoint'] = out[2].detach().numpy()
  t['pred_semantic'] = sub(out[3])
  grads = dict(net.named_parameters())
  for k in ('backbone.image_encoder.stem.conv.weight', 'backbone.lidar_encoder.stem.conv.weight',
            'backbone.image_encoder.s1.b1.conv2.conv.weight', 'backbone.image_encoder.s3.b7.conv1.conv.weight',
            'backbone.lidar_encoder.s4.b1.se.fc1.weight', 'backbone.transformers.0.blocks.0.attn.query.weight',
            'backbone.transformers.3.blocks.1.mlp.0.bias', 'backbone.transformers.2.pos_emb',
            'backbone.lidar_channel_to_img.1.weight', 'backbone.up_conv4.weight', 'change_channel.weight',
            'join.layers.0.self_attn.in_proj_weight', 'join.layers.5.linear2.weight',
            'checkpoint_decoder.gru.weight_hh_l0', 'checkpoint_decoder.encoder.weight', 'checkpoint_query',
            'target_speed_network.2.weight', 'semantic_decoder.deconv1.0.weight', 'semantic_decoder.deconv3.2.weight',
            'depth_decoder.deconv2.0.weight', 'bev_semantic_decoder.2.weight', 'head.heatmap_head.0.weight',
            'head.yaw_class_head.2.bias', 'extra_sensor_encoder.0.weight',
            'backbone.image_encoder.s2.b3.conv3.bn.weight', 'backbone.lidar_encoder.s3.b2.conv2.bn.bias'):
    gr = grads[k].grad
    t['gradnorm_' + k] = np.array(float(gr.norm()))
    t['grad_' + k] = gr.flatten()[:256].numpy().copy()
  np.savez_compressed(os.path.join(OUT, 'train_b2.npz'), **t)
  for f in sorted(os.listdir(OUT)):
    print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
  main()
