"""Drop-in for team_code/model.py: put this directory on PYTHONPATH ahead of /path/to/carla_garage/team_code and
`from model import LidarCenterNet` (train.py:25, sensor_agent.py:20) resolves to the B200-native class."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from carla_garage_b200.nn.model import (LidarCenterNet, GRUWaypointsPredictorInterFuser,  # noqa: E402,F401
                                        GRUWaypointsPredictorTransFuser, PositionEmbeddingSine)
