"""Generates tests/golden/bev_b2.npz + bev_keys.json by running the UNMODIFIED reference (/root/reference/team_code)
with ``config.backbone = 'bev_encoder'`` (bev_encoder.py: SimpleBEV-style lift of the camera features into the BEV grid,
concatenated with the LiDAR histogram, one RegNet over the fused BEV).

  python tests/golden/make_golden_bev.py

Stored: eval-mode outputs + taps, train-mode outputs / the ten losses / slices + norms of EVERY parameter gradient, and
fingerprints of the three geometry buffers (grid, normaliser, visibility mask) the framework recomputes itself."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from carla_garage_b200 import compat, synth  # noqa: E402
from oracle.regnety import timm_factory  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
GEOMETRY = ('backbone.grid', 'backbone.bev_projection_normalizer', 'backbone.valid_bev_pixels')


def main():
  compat.install(timm_factory)
  from config import GlobalConfig  # pylint: disable=import-outside-toplevel
  from model import LidarCenterNet  # pylint: disable=import-outside-toplevel
  torch.manual_seed(0)
  torch.set_num_threads(min(os.cpu_count() or 1, 16))
  cfg = GlobalConfig()
  cfg.backbone = 'bev_encoder'
  net = LidarCenterNet(cfg)
  ref_sd = net.state_dict()
  shapes = {k: list(v.shape) for k, v in ref_sd.items()}
  json.dump(dict(shapes=shapes, order=list(ref_sd.keys())), open(os.path.join(OUT, 'bev_keys.json'), 'w'), indent=0)
  fixed = {k: ref_sd[k] for k in ('valid_bev_pixels', 'valid_bev_pixels_inv', 'loss_speed.weight', 'loss_semantic.weight',
                                  'loss_bev_semantic.weight') + GEOMETRY}
  sd = synth.make_state_dict(shapes, seed=0, fixed=fixed)
  net.load_state_dict(sd, strict=True)
  for m in net.modules():
    if isinstance(m, torch.nn.Dropout):
      m.p = 0.0
    if isinstance(m, torch.nn.MultiheadAttention):
      m.dropout = 0.0
  B = 2
  inp = synth.make_inputs(B, seed=11)
  lab = synth.make_labels(B, seed=13)
  g = {}
  # geometry fingerprints
  grid = ref_sd['backbone.grid']
  g['grid_shape'] = np.array(grid.shape)
  g['grid_sample'] = grid.flatten()[::997].numpy().copy()
  g['grid_sum'] = np.array(float(grid.double().sum()))
  g['normalizer'] = ref_sd['backbone.bev_projection_normalizer'].numpy().copy().astype(np.float32)
  g['backbone_valid_bev_pixels'] = np.packbits(ref_sd['backbone.valid_bev_pixels'].numpy().astype(np.uint8))

  # eval forward + taps
  taps = {}
  hooks = []

  def tap(name, mod):
    hooks.append(mod.register_forward_hook(lambda _m, _i, o: taps.__setitem__(name, o.detach().clone())))

  bb = net.backbone
  tap('upsampled', bb.upsampling_layer)
  tap('image_features', bb.depth_layer)
  tap('bev_compressed', bb.bev_compressor)
  tap('bev_s1', bb.bev_encoder.s1)
  tap('bev_s3', bb.bev_encoder.s3)
  net.eval()
  with torch.no_grad():
    out = net(**inp)
  for h in hooks:
    h.remove()
  for k, v in taps.items():
    g['eval_tap_' + k] = v[:, :8, ::2, ::2].numpy().copy() if v.dim() == 4 else v.numpy()
    g['eval_tapnorm_' + k] = np.array(float(v.double().norm()))
  names = ('pred_wp', 'pred_target_speed', 'pred_checkpoint', 'pred_semantic', 'pred_bev_semantic', 'pred_depth')
  for n, o in zip(names, out[:6]):
    if o is not None:
      g['eval_' + n] = o[..., ::4, ::4].numpy().copy() if o.numel() > 1e5 else o.numpy()  # big maps: every 4th pixel + norm
      g['evalnorm_' + n] = np.array(float(o.double().norm()))
  for n, o in zip(('heatmap', 'wh', 'offset', 'yaw_class', 'yaw_res'), out[6][:5]):
    g['eval_box_' + n] = o.numpy()

  # train step
  net.train()
  out = net(**inp)
  losses = net.compute_loss(pred_wp=out[0], pred_target_speed=out[1], pred_checkpoint=out[2], pred_semantic=out[3],
                            pred_bev_semantic=out[4], pred_depth=out[5], pred_bounding_box=out[6], pred_wp_1=out[8],
                            selected_path=out[9], waypoint_label=None, target_speed_label=lab['target_speed'],
                            checkpoint_label=lab['checkpoint'], semantic_label=lab['semantic'],
                            bev_semantic_label=lab['bev_semantic'], depth_label=lab['depth'],
                            center_heatmap_label=lab['center_heatmap'], wh_label=lab['wh'],
                            yaw_class_label=lab['yaw_class'], yaw_res_label=lab['yaw_res'], offset_label=lab['offset'],
                            velocity_label=None, brake_target_label=None, pixel_weight_label=lab['pixel_weight'],
                            avg_factor_label=lab['avg_factor'])
  total = sum(losses.values()) / len(losses)
  total.backward()
  for k, v in losses.items():
    g[k] = np.array(float(v))
  g['total'] = np.array(float(total))
  g['train_pred_checkpoint'] = out[2].detach().numpy()
  g['train_pred_target_speed'] = out[1].detach().numpy()
  for n, p in net.named_parameters():
    if p.grad is not None:
      g['grad_' + n] = p.grad.flatten()[:256].numpy().copy()
      g['gradnorm_' + n] = np.array(float(p.grad.double().norm()))
  np.savez_compressed(os.path.join(OUT, 'bev_b2.npz'), **g)
  print(len(g), 'entries', os.path.getsize(os.path.join(OUT, 'bev_b2.npz')) / 1e6, 'MB')
  print({k: float(v) for k, v in losses.items()})


if __name__ == '__main__':
  main()
