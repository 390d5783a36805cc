"""``Box2BoxXYXYTransform`` with the reference's interface (ubteacher/modeling/box_regression.py:12-129) on the device:
deltas (dl, dr, dd, du) = w * (target - src) / (src_size + 1); apply: pred = d / w * src_size (NO + 1) + coord, clamped to
+-SCALE_CLAMP. Bit-identical to the reference's torch arithmetic (csrc/roi.cu: round-to-nearest intrinsics, same order).
The training / inference kernels apply the same formulas in registers; this is the stand-alone operator."""
import torch

from .. import _C
from .._C import f32

_DEFAULT_SCALE_CLAMP = 1000.0 / 16


class Box2BoxXYXYTransform:
    def __init__(self, weights, scale_clamp=_DEFAULT_SCALE_CLAMP):
        self.weights = tuple(float(w) for w in weights)
        self.scale_clamp = float(scale_clamp)

    def get_deltas(self, src_boxes, target_boxes):
        assert isinstance(src_boxes, torch.Tensor), type(src_boxes)
        assert isinstance(target_boxes, torch.Tensor), type(target_boxes)
        if not src_boxes.is_cuda:
            raise RuntimeError("Box2BoxXYXYTransform runs on the device (no CPU fallback)")
        src = src_boxes.detach().float().contiguous()
        tgt = target_boxes.detach().float().contiguous()
        out = torch.empty_like(src)
        _C.counted_call("ut2_box2box_xyxy_get_deltas", src, tgt, src.shape[0], f32(self.weights[0]), f32(self.weights[1]), out)
        return out

    def apply_deltas(self, deltas, boxes):
        if not deltas.is_cuda:
            raise RuntimeError("Box2BoxXYXYTransform runs on the device (no CPU fallback)")
        d = deltas.detach().float().contiguous()
        b = boxes.detach().float().contiguous()
        n, k = d.shape[0], d.shape[1] // 4
        out = torch.empty_like(d)
        _C.counted_call("ut2_box2box_xyxy_apply_deltas", d, b, n, k, f32(self.weights[0]), f32(self.weights[1]),
                        f32(self.scale_clamp), out)
        return out
