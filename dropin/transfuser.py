"""Drop-in for team_code/transfuser.py (`from transfuser import TransfuserBackbone`, model.py:8)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from carla_garage_b200.nn.transfuser import TransfuserBackbone, GPT, Block, SelfAttention  # noqa: E402,F401
