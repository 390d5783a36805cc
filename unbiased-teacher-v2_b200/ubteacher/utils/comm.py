"""ubteacher.utils.comm (reference: ubteacher/utils/comm.py:7-45): reduce_sum and compute_locations."""
import torch

from ..d2compat.comm import get_world_size, reduce_sum  # noqa: F401


def compute_locations(h, w, stride, device):
    """(x, y) = stride * index + stride // 2, row-major over (h, w)  — utils/comm.py:34-45."""
    ys = torch.arange(0, h * stride, step=stride, dtype=torch.float32, device=device)
    xs = torch.arange(0, w * stride, step=stride, dtype=torch.float32, device=device)
    return torch.stack((xs.repeat(h), ys.repeat_interleave(w)), dim=1) + stride // 2
