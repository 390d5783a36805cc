"""``ubteacher.layers`` with the reference's names (ubteacher/layers/__init__.py): IOULoss, KLLoss, NLLoss, ml_nms — device
operators behind the module interface; the training step uses the fused kernels (csrc/fcos_loss.cu, csrc/proposals.cu)."""
from .iou_loss import IOULoss  # noqa: F401
from .kl_loss import KLLoss, NLLoss  # noqa: F401
from .ml_nms import ml_nms  # noqa: F401

__all__ = ["IOULoss", "KLLoss", "NLLoss", "ml_nms"]
