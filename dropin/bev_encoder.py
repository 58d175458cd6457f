"""Drop-in for team_code/bev_encoder.py (`from bev_encoder import BevEncoder`, model.py:9)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from carla_garage_b200.nn.bev_encoder import BevEncoder, UpsamplingConcat  # noqa: E402,F401
