"""Generates tests/golden/mlp_join_b2.npz + mlp_join_keys.json by running the UNMODIFIED reference
(/root/reference/team_code) with ``transformer_decoder_join = False`` and ``use_wp_gru = True`` — the original TransFuser
planner: global-pool MLP join + GRUWaypointsPredictorTransFuser heads (model.py:184-209,359-376,870-913).

  python tests/golden/make_golden_mlp_join.py
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

OUT = os.path.join(ROOT, 'tests', 'golden')
PLANNER = ('join.', 'wp_decoder.', 'checkpoint_decoder.', 'target_speed_network.', 'extra_sensor_encoder.',
           'backbone.lidar_to_img_features_end.')


def main():
  compat.install(timm_factory)
  from config import GlobalConfig  # pylint: disable=import-outside-toplevel
  from model import LidarCenterNet  # pylint: disable=import-outside-toplevel
  torch.manual_seed(0)
  torch.set_num_threads(min(os.cpu_count() or 1, 16))
  cfg = GlobalConfig()
  cfg.transformer_decoder_join = False
  cfg.use_wp_gru = True
  net = LidarCenterNet(cfg)
  ref_sd = net.state_dict()
  shapes = {k: list(v.shape) for k, v in ref_sd.items()}
  json.dump(dict(shapes=shapes, order=list(ref_sd.keys())), open(os.path.join(OUT, 'mlp_join_keys.json'), 'w'), indent=0)
  fixed = {k: ref_sd[k] for k in ('valid_bev_pixels', 'valid_bev_pixels_inv', 'loss_speed.weight', 'loss_semantic.weight',
                                  'loss_bev_semantic.weight')}
  sd = synth.mlp_join_tweak(synth.make_state_dict(shapes, seed=0, fixed=fixed))   # keeps a ReLU pre-activation off 0
  net.load_state_dict(sd, strict=True)
  net.train()
  for m in net.modules():
    if isinstance(m, torch.nn.Dropout):
      m.p = 0.0
  B = 2
  inp = synth.make_inputs(B, seed=11)
  lab = synth.make_labels(B, seed=13)
  wp_lab = synth.make_waypoint_labels(B, cfg.pred_len // cfg.wp_dilation, seed=13)
  out = net(**inp)
  losses = net.compute_loss(pred_wp=out[0], pred_target_speed=out[1], pred_checkpoint=out[2], pred_semantic=out[3],
                            pred_bev_semantic=out[4], pred_depth=out[5], pred_bounding_box=out[6], pred_wp_1=out[8],
                            selected_path=out[9], waypoint_label=wp_lab, target_speed_label=lab['target_speed'],
                            checkpoint_label=lab['checkpoint'], semantic_label=lab['semantic'],
                            bev_semantic_label=lab['bev_semantic'], depth_label=lab['depth'],
                            center_heatmap_label=lab['center_heatmap'], wh_label=lab['wh'],
                            yaw_class_label=lab['yaw_class'], yaw_res_label=lab['yaw_res'], offset_label=lab['offset'],
                            velocity_label=None, brake_target_label=None, pixel_weight_label=lab['pixel_weight'],
                            avg_factor_label=lab['avg_factor'])
  total = sum(losses.values()) / len(losses)
  total.backward()
  g = {k: np.array(float(v)) for k, v in losses.items()}
  g['total'] = np.array(float(total))
  g['pred_wp'] = out[0].detach().numpy()
  g['pred_target_speed'] = out[1].detach().numpy()
  g['pred_checkpoint'] = out[2].detach().numpy()
  for n, p in net.named_parameters():
    if n.startswith(PLANNER):
      g['grad_' + n] = p.grad.flatten()[:512].numpy().copy()
      g['gradnorm_' + n] = np.array(float(p.grad.norm()))
  np.savez_compressed(os.path.join(OUT, 'mlp_join_b2.npz'), **g)
  print(sorted(k for k in g if not k.startswith('grad')), os.path.getsize(os.path.join(OUT, 'mlp_join_b2.npz')))


if __name__ == '__main__':
  main()
