"""Autograd plumbing shared by the loss layers: the kernels return the value and the input gradients of one call; backward scales
the stored gradients by the incoming scalar gradient on the device (ut2_scale_f32)."""
import torch

from .. import _C
from .._C import i64


def _rows(t):
    if not t.is_cuda:
        raise RuntimeError("ubteacher.layers run on the device (no CPU fallback)")
    return t.detach().float().contiguous()


def scale(g, s):
    out = torch.empty_like(g)
    _C.counted_call("ut2_scale_f32", g, s.detach().float().reshape(1).contiguous(), out, i64(g.numel()))
    return out


class ScalarLoss(torch.autograd.Function):
    """forward(ctx, launch, *inputs): `launch(inputs, needs)` -> (loss[1], [grad or None per input]); needs = which inputs want one."""

    @staticmethod
    def forward(ctx, launch, *inputs):
        loss, grads = launch(inputs, ctx.needs_input_grad[1:])
        ctx.grads = grads
        ctx.dtypes = [t.dtype if isinstance(t, torch.Tensor) else None for t in inputs]
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        out = [None]
        for g, dt in zip(ctx.grads, ctx.dtypes):
            out.append(None if g is None else scale(g, gout).to(dt))
        return tuple(out)
