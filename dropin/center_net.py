"""Drop-in for team_code/center_net.py (`from center_net import LidarCenterNetHead`, model.py:12;
`from center_net import angle2class`, data.py:21)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from carla_garage_b200.nn.center_net import LidarCenterNetHead  # noqa: E402,F401


def angle2class(angle, num_dir_bins):
  """center_net.py:240-254 (host-side label helper used by the data loader)."""
  angle = angle % (2 * np.pi)
  angle_per_class = 2 * np.pi / float(num_dir_bins)
  shifted_angle = (angle + angle_per_class / 2) % (2 * np.pi)
  angle_cls = shifted_angle // angle_per_class
  angle_res = shifted_angle - (angle_cls * angle_per_class + angle_per_class / 2)
  return int(angle_cls), angle_res
