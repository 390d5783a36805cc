"""``ubteacher.layers`` with the reference's names (ubteacher/layers/__init__.py): IOULoss, KLLoss, NLLoss, ml_nms — device
operators behind the module interface; the training step uses the fused kernels (csrc/fcos_loss.cu, csrc/proposals.cu)."""
from . import iou_loss as _iou, kl_loss as _kl, ml_nms as _nms

IOULoss = _iou.IOULoss
KLLoss, NLLoss = _kl.KLLoss, _kl.NLLoss
ml_nms = _nms.ml_nms

__all__ = ["IOULoss", "KLLoss", "NLLoss", "ml_nms"]
