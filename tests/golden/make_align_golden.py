"""Generates tests/golden/align.npz by running the UNMODIFIED reference: CARLA_Data.align (data.py:840-871) followed by
CARLA_Data.lidar_to_histogram_features (data.py:873-906) on seeded clouds and ego poses.

  python tests/golden/make_align_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from carla_garage_b200 import compat, synth  # noqa: E402
from oracle.regnety import timm_factory  # noqa: E402

CASES = [  # (pos0, theta0, pos1, theta1, y_aug, yaw_aug_deg)
    ((10.0, -4.0), 0.3, (11.5, -3.2), 0.42, 0.0, 0.0),
    ((-120.25, 33.0), 3.0, (-121.0, 34.5), -3.1, 0.7, -12.0),
    ((0.0, 0.0), 0.0, (0.0, 0.0), 0.0, 0.0, 0.0),           # identity: must equal the unaligned histogram of float64 points
    ((5.0, 5.0), -1.2, (3.0, 9.0), -0.8, -1.0, 20.0),
]


def main():
  compat.install(timm_factory)
  from config import GlobalConfig  # pylint: disable=import-outside-toplevel
  from data import CARLA_Data  # pylint: disable=import-outside-toplevel
  data = CARLA_Data(root=[], config=GlobalConfig(), shared_dict=None)
  pts = synth.make_point_clouds(len(CASES), seed=21, n_points=20000).numpy()
  out = {'cases': np.array([[c[0][0], c[0][1], c[1], c[2][0], c[2][1], c[3], c[4], c[5]] for c in CASES], np.float64)}
  for i, (p0, t0, p1, t1, ya, yw) in enumerate(CASES):
    m0, m1 = {'pos_global': p0, 'theta': t0}, {'pos_global': p1, 'theta': t1}
    aligned = data.align(pts[i], m0, m1, y_augmentation=ya, yaw_augmentation=yw)
    for gp in (False, True):
      h = data.lidar_to_histogram_features(aligned, use_ground_plane=gp)
      out[f'hist{i}_gp{int(gp)}'] = np.round(h * 5).astype(np.uint8)
  np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'align.npz'), **out)
  print(os.path.getsize(os.path.join(ROOT, 'tests', 'golden', 'align.npz')), 'bytes')


if __name__ == '__main__':
  main()
