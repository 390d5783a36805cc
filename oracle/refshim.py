"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

Loads selected files of the *reference* (``/root/reference/ubteacher/...``) by path, in this
container only, so that golden fixtures can be minted from the reference's own arithmetic
(SURVEY.md F5). Detectron2 / fvcore are not installed, so the handful of symbols those files import
are stubbed: containers come from the product's d2compat shims, ``batched_nms`` from torchvision,
``sigmoid_focal_loss_jit`` / ``smooth_l1_loss`` / ``giou_loss`` from their published formulas
(SURVEY.md B.5), and ``torch.Tensor.cuda`` is patched to the identity because the reference
hard-codes ``.cuda()``. Nothing from the reference is copied into this repo.
"""
import importlib.util
import os
import sys
import types

import torch

REF = os.environ.get("UT2_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "unbiased-teacher-v2_b200")


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _focal(inputs, targets, alpha=-1, gamma=2, reduction="none"):
    import torch.nn.functional as F

    p = torch.sigmoid(inputs)
    ce = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    p_t = p * targets + (1 - p) * (1 - targets)
    loss = ce * ((1 - p_t) ** gamma)
    if alpha >= 0:
        loss = (alpha * targets + (1 - alpha) * (1 - targets)) * loss
    if reduction == "mean":
        loss = loss.mean()
    elif reduction == "sum":
        loss = loss.sum()
    return loss


def _smooth_l1(input, target, beta, reduction="none"):
    if beta < 1e-5:
        loss = torch.abs(input - target)
    else:
        n = torch.abs(input - target)
        loss = torch.where(n < beta, 0.5 * n**2 / beta, n - 0.5 * beta)
    if reduction == "mean":
        loss = loss.mean() if loss.numel() > 0 else 0.0 * loss.sum()
    elif reduction == "sum":
        loss = loss.sum()
    return loss


def install(world_size=1):
    """Install the stubs and return a loader ``load(relpath, modname)``."""
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference tree {REF} not present (fixtures can only be minted in the build container)")
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    # our package is also called `ubteacher`; load its compat module by path to avoid the clash
    spec = importlib.util.spec_from_file_location(
        "_ut2_structures", os.path.join(PKG, "ubteacher", "d2compat", "structures.py"))
    st = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(st)
    import torchvision

    torch.Tensor.cuda = lambda self, *a, **k: self  # reference hard-codes .cuda()

    def batched_nms(boxes, scores, idxs, iou_threshold):
        return torchvision.ops.boxes.batched_nms(boxes.float(), scores, idxs, iou_threshold)

    def nonzero_tuple(x):
        if x.dim() == 0:
            return x.unsqueeze(0).nonzero().unbind(1)
        return x.nonzero().unbind(1)

    def configurable(f=None, **kw):
        return f if f is not None else (lambda g: g)

    _mod("detectron2")
    _mod("detectron2.layers", cat=st.cat, batched_nms=batched_nms, ShapeSpec=st.ShapeSpec,
         nonzero_tuple=nonzero_tuple, cross_entropy=torch.nn.functional.cross_entropy)
    _mod("detectron2.structures", Boxes=st.Boxes, Instances=st.Instances, ImageList=st.ImageList)
    _mod("detectron2.structures.instances", Instances=st.Instances)
    _mod("detectron2.utils")
    _mod("detectron2.utils.comm", get_world_size=lambda: world_size)
    _mod("detectron2.config", configurable=configurable)
    _mod("detectron2.modeling")
    _mod("detectron2.modeling.roi_heads")
    _mod("detectron2.modeling.roi_heads.fast_rcnn", FastRCNNOutputLayers=torch.nn.Module,
         fast_rcnn_inference=None, _log_classification_stats=lambda *a, **k: None)
    _mod("detectron2.modeling.box_regression", Box2BoxTransform=object)
    _mod("detectron2.utils.events", get_event_storage=lambda: None)
    _mod("fvcore")
    _mod("fvcore.nn", sigmoid_focal_loss_jit=_focal, smooth_l1_loss=_smooth_l1, giou_loss=None)

    # empty parent packages whose __path__ points at the reference tree (their __init__ is NOT run)
    for name, rel in [("ubteacher", "ubteacher"), ("ubteacher.layers", "ubteacher/layers"),
                      ("ubteacher.utils", "ubteacher/utils"), ("ubteacher.modeling", "ubteacher/modeling"),
                      ("ubteacher.modeling.fcos", "ubteacher/modeling/fcos"),
                      ("ubteacher.modeling.roi_heads", "ubteacher/modeling/roi_heads")]:
        m = _mod(name)
        m.__path__ = [os.path.join(REF, rel)]

    def load(relpath, modname):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
        m = importlib.util.module_from_spec(spec)
        sys.modules[modname] = m
        spec.loader.exec_module(m)
        return m

    iou = load("ubteacher/layers/iou_loss.py", "ubteacher.layers.iou_loss")
    kl = load("ubteacher/layers/kl_loss.py", "ubteacher.layers.kl_loss")
    nms = load("ubteacher/layers/ml_nms.py", "ubteacher.layers.ml_nms")
    lay = sys.modules["ubteacher.layers"]
    lay.IOULoss, lay.KLLoss, lay.NLLoss, lay.ml_nms = iou.IOULoss, kl.KLLoss, kl.NLLoss, nms.ml_nms
    load("ubteacher/utils/comm.py", "ubteacher.utils.comm")
    return load, st
