from .one_stage_detector import OneStageDetector, PseudoProposalNetwork  # noqa: F401
from .pseudo_generator import PseudoGenerator  # noqa: F401
from .meta_arch.rcnn import TwoStagePseudoLabGeneralizedRCNN  # noqa: F401
