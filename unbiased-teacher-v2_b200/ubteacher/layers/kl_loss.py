"""``KLLoss`` / ``NLLoss`` (ubteacher/layers/kl_loss.py:11-105). NLLoss is the uncertainty-aware regression term of the shipped
recipes (KL_LOSS_TYPE "nlloss"): mean over boxes of [sum_4 ((t - mu)^2 / (2 sigma^2) + 0.5 log sigma^2) + 2 log(2 pi)] * iou_weight,
sigma = sigmoid(input_std) — ``weight`` / ``beta`` / ``loss_denorm`` / ``method`` are accepted and ignored like in the reference.
KLLoss: exp(-std) * smooth_l1 + 0.5 * std with the four reductions; ``beta < 1e-5`` returns None in the reference (its
reductions sit inside the else branch) and raises here. Device operators ``ut2_nl_loss`` / ``ut2_kl_loss``; differentiable
w.r.t. ``input`` and ``input_std``."""
import torch
from torch import nn

from .. import _C
from .._C import f32
from ._fn import ScalarLoss, _rows

_METHODS = {"weight_ctr_sum": 0, "weight_ctr_mean": 1, "sum": 2, "mean": 3}


class NLLoss(nn.Module):
    def forward(self, input, input_std, target, weight=None, iou_weight=None, beta=1.0, loss_denorm=None, method="weight_ctr_sum"):
        assert iou_weight is not None, "NLLoss multiplies by iou_weight (kl_loss.py:103)"

        def launch(inputs, needs):
            m, s, t, w = _rows(inputs[0]), _rows(inputs[1]), _rows(target), _rows(iou_weight)
            acc = torch.empty(1, dtype=torch.float64, device=m.device)
            loss = torch.empty(1, dtype=torch.float32, device=m.device)
            dm = torch.empty_like(m) if needs[0] else None
            ds = torch.empty_like(s) if needs[1] else None
            _C.counted_call("ut2_nl_loss", m, s, t, w, m.shape[0], acc, loss, dm, ds)
            return loss, [dm, ds]
        return ScalarLoss.apply(launch, input, input_std)


class KLLoss(nn.Module):
    def forward(self, input, input_std, target, weight=None, iou_weight=None, beta=1.0, loss_denorm=None, method="weight_ctr_sum"):
        if beta < 1e-5:
            raise NotImplementedError("KLLoss with beta < 1e-5: the reference computes |input - target| and returns None")
        if method not in _METHODS:
            raise ValueError("No defined regression loss method")
        meth = _METHODS[method]
        if meth < 2:
            assert weight is not None
        if meth == 1:
            assert loss_denorm is not None

        def launch(inputs, needs):
            x, s, t = _rows(inputs[0]), _rows(inputs[1]), _rows(target)
            w = _rows(weight) if weight is not None else None
            acc = torch.empty(1, dtype=torch.float64, device=x.device)
            loss = torch.empty(1, dtype=torch.float32, device=x.device)
            dx = torch.empty_like(x) if needs[0] else None
            ds = torch.empty_like(s) if needs[1] else None
            _C.counted_call("ut2_kl_loss", x, s, t, w, x.shape[0], f32(beta), meth, f32(float(loss_denorm) if loss_denorm is not None else 1.0),
                            acc, loss, dx, ds)
            return loss, [dx, ds]
        return ScalarLoss.apply(launch, input, input_std)
