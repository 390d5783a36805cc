from .config import CfgNode, get_cfg  # noqa: F401
from .structures import Boxes, ImageList, Instances, ShapeSpec, cat  # noqa: F401
