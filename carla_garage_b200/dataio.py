"""Host-side helpers of the input pipeline pieces that run on the device: the rigid alignment that
CARLA_Data.align (team_code/data.py:840-871) applies to a past LiDAR sweep becomes two {translation, yaw} transforms
handed to the pillar-scatter kernel (tfpp_pillar_scatter_aligned), which applies them per point in float64."""
import math

import numpy as np


def normalize_angle(x):
  """transfuser_utils.normalize_angle (transfuser_utils.py:57-63): wrap to (-pi, pi]."""
  x = x % (2 * np.pi)
  if x > np.pi:
    x -= 2 * np.pi
  return x


def align_transforms(measurements_0, measurements_1, y_augmentation=0.0, yaw_augmentation=0.0):
  """The two algin_lidar calls of data.py:855-869 as a (2, 4) float64 array {tx, ty, tz, yaw}: ego motion from the frame
  of ``measurements_0`` to the frame of ``measurements_1`` (dicts with 'pos_global' and 'theta'), then the augmentation
  shift / rotation.  p' = R(yaw)^T (p - t) each (transfuser_utils.py:116-130)."""
  pos_1 = np.array([measurements_1['pos_global'][0], measurements_1['pos_global'][1], 0.0])
  pos_0 = np.array([measurements_0['pos_global'][0], measurements_0['pos_global'][1], 0.0])
  pos_diff = pos_1 - pos_0
  rot_diff = normalize_angle(measurements_1['theta'] - measurements_0['theta'])
  th = measurements_1['theta']
  rotation_matrix = np.array([[np.cos(th), -np.sin(th), 0.0], [np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]])
  pos_diff = rotation_matrix.T @ pos_diff
  return np.array([[pos_diff[0], pos_diff[1], pos_diff[2], rot_diff],
                   [0.0, y_augmentation, 0.0, math.radians(yaw_augmentation)]], dtype=np.float64)
