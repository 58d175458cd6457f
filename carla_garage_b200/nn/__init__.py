from .model import (LidarCenterNet, GRUWaypointsPredictorInterFuser, GRUWaypointsPredictorTransFuser,  # noqa: F401
                    PositionEmbeddingSine)
from .transfuser import TransfuserBackbone, GPT  # noqa: F401
from .center_net import LidarCenterNetHead  # noqa: F401
