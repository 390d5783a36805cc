from .fast_rcnn import FastRCNNFocaltLossBoundaryVarOutputLayers  # noqa: F401
from .roi_heads import StandardROIHeadsPseudoLab  # noqa: F401
