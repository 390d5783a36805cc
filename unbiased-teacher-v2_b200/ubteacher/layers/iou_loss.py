"""``IOULoss`` (ubteacher/layers/iou_loss.py:9-76): IoU / linear IoU / gIoU loss on (l, t, r, b) distance rows, summed (optionally
weighted). Device operator ``ut2_iou_loss``; differentiable w.r.t. ``pred``."""
import torch
from torch import nn

from .. import _C
from ._fn import ScalarLoss, _rows

_TYPES = {"iou": 0, "linear_iou": 1, "giou": 2}


class IOULoss(nn.Module):
    def __init__(self, loc_loss_type="iou"):
        super().__init__()
        self.loc_loss_type = loc_loss_type

    def forward(self, pred, target, weight=None):
        if self.loc_loss_type not in _TYPES:
            raise NotImplementedError
        kind = _TYPES[self.loc_loss_type]

        def launch(inputs, needs):
            p, t = _rows(inputs[0]), _rows(target)
            w = _rows(weight) if weight is not None else None
            P = p.shape[0]
            acc = torch.empty(1, dtype=torch.float64, device=p.device)
            loss = torch.empty(1, dtype=torch.float32, device=p.device)
            dp = torch.empty_like(p) if needs[0] else None
            _C.counted_call("ut2_iou_loss", p, t, w, P, kind, acc, loss, dp)
            return loss, [dp]
        return ScalarLoss.apply(launch, pred)
