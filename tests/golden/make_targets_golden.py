"""Generates tests/golden/targets.npz by running the UNMODIFIED reference rasteriser
(team_code/data.py:698-791 CARLA_Data.get_targets + gaussian_target.py) in the build container.

  python tests/golden/make_targets_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from carla_garage_b200 import compat, synth  # noqa: E402
from oracle.regnety import timm_factory  # noqa: E402


def main():
  compat.install(timm_factory)
  from config import GlobalConfig  # pylint: disable=import-outside-toplevel
  from data import CARLA_Data  # pylint: disable=import-outside-toplevel
  cfg = GlobalConfig()
  data = CARLA_Data(root=[], config=cfg, shared_dict=None)
  out = {}
  cases = synth.make_gt_boxes(6, seed=3)
  cases.append(np.zeros((0, 8), np.float32))                                        # no boxes
  cases.append(np.array([[0.5, 0.5, 9.0, 4.0, 0.1, 1.0, 0.4, 0], [255.9, 255.9, 6.0, 12.0, -3.0, 2.0, 0.6, 3],
                         [130.2, 0.0, 30.0, 30.0, 3.14159, 0.0, 0.5, 1], [130.9, 3.9, 2.0, 2.0, 6.9, 0.0, 1.0, 1],
                         [131.5, 2.1, 5.0, 7.0, -0.26, 3.0, 0.0, 2]], np.float32))  # borders, one shared centre pixel
  for i, boxes in enumerate(cases):
    t, avg = data.get_targets(boxes, cfg.lidar_resolution_height // cfg.bev_down_sample_factor,
                              cfg.lidar_resolution_width // cfg.bev_down_sample_factor)
    out[f'boxes{i}'] = boxes
    out[f'avg{i}'] = np.array(avg)
    for k, v in t.items():
      out[f'{k}{i}'] = np.asarray(v)
  np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'targets.npz'), **out)
  print(len(cases), 'cases', os.path.getsize(os.path.join(ROOT, 'tests', 'golden', 'targets.npz')), 'bytes')


if __name__ == '__main__':
  main()
