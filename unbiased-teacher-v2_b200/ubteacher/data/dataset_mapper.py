"""``DatasetMapperTwoCropSeparate`` for decoded images, on the device (reference: ubteacher/data/dataset_mapper.py:73-157).

The reference mapper reads an image file, applies Detectron2's weak augmentation ([D2] ResizeShortestEdge + RandomFlip,
`utils.build_augmentation`), transforms the annotations, applies the torchvision strong augmentation to a PIL copy and
returns two dicts: (strong view with the labels, weak view with the same labels). Here everything after the file decode
runs in CUDA kernels (csrc/augment.cu): Pillow's bilinear resample + flip (`ut2_resize_flip_u8`, bit-exact) and the strong
augmentation (`ut2_strong_augment_u8`, bit-exact); the box arithmetic (a few numbers per image) stays on the host.

Random draws follow the reference: numpy's global RNG for the resize size and the flip ([D2] transforms), torch / `random`
for the strong augmentation (torchvision). Detectron2 is not on disk: its parameter logic is restated."""
import ctypes

import numpy as np
import torch

from .. import _C
from .._C import i64
from ..d2compat.structures import Boxes, Instances
from .gpu_augmentation import GpuStrongAugmentation


def shortest_edge_shape(h, w, size, max_size):
    """[D2] ResizeShortestEdge.get_output_shape."""
    scale = size * 1.0 / min(h, w)
    if h < w:
        newh, neww = size, scale * w
    else:
        newh, neww = scale * h, size
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh, neww = newh * scale, neww * scale
    return int(newh + 0.5), int(neww + 0.5)


def resize_flip(img_hwc, new_h, new_w, flip):
    """uint8 [h, w, 3] CUDA tensor -> uint8 [3, new_h, new_w] (Pillow BILINEAR resize, optional horizontal flip)."""
    assert img_hwc.is_cuda and img_hwc.dtype == torch.uint8 and img_hwc.dim() == 3 and img_hwc.shape[2] == 3 and img_hwc.is_contiguous()
    h, w = int(img_hwc.shape[0]), int(img_hwc.shape[1])
    dev = img_hwc.device
    out = torch.empty((3, new_h, new_w), dtype=torch.uint8, device=dev)
    tmp = torch.empty((h, new_w, 3), dtype=torch.uint8, device=dev)
    fn = _C.lib().ut2_resize_workspace_bytes
    fn.restype = ctypes.c_longlong
    wsb = fn(h, w, new_h, new_w)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    _C.counted_call("ut2_resize_flip_u8", img_hwc, h, w, out, new_h, new_w, int(bool(flip)), tmp, ws, i64(wsb))
    _C.launch_count += 3
    return out


def transform_boxes(boxes, h, w, new_h, new_w, flip):
    """[D2] transform_instance_annotations (XYXY_ABS) + filter_empty_instances. Returns (float32 [n, 4], keep mask)."""
    b = np.asarray(boxes, dtype=np.float64).reshape(-1, 4)
    sx, sy = new_w * 1.0 / w, new_h * 1.0 / h
    x0, y0, x1, y1 = b[:, 0] * sx, b[:, 1] * sy, b[:, 2] * sx, b[:, 3] * sy
    if flip:
        x0, x1 = new_w - x1, new_w - x0
    out = np.stack([np.minimum(x0, x1), np.minimum(y0, y1), np.maximum(x0, x1), np.maximum(y0, y1)], axis=1).clip(min=0)
    out = np.minimum(out, np.array([new_w, new_h, new_w, new_h], dtype=np.float64))
    keep = ((out[:, 2] - out[:, 0]) > 1e-5) & ((out[:, 3] - out[:, 1]) > 1e-5)
    return out.astype(np.float32), keep


class DatasetMapperTwoCropSeparate:
    """``mapper(list of dataset dicts) -> (list of strong-view dicts, list of weak-view dicts)``. A dataset dict holds the
    decoded image as ``"image"`` (uint8 [h, w, 3], numpy / CPU / CUDA tensor, in cfg.INPUT.FORMAT channel order) and
    optionally ``"annotations"`` = [{"bbox": xyxy abs, "category_id": int, "iscrowd": 0|1}] (Detectron2 dataset format,
    XYXY_ABS boxes). Batched because the device kernels are: one call per training batch."""

    def __init__(self, cfg, is_train=True):
        self.is_train = is_train
        inp = cfg.INPUT
        self.min_size = tuple(inp.MIN_SIZE_TRAIN if is_train else (inp.MIN_SIZE_TEST,))
        self.max_size = inp.MAX_SIZE_TRAIN if is_train else inp.MAX_SIZE_TEST
        self.sample_style = inp.MIN_SIZE_TRAIN_SAMPLING if is_train else "choice"
        if self.sample_style == "range":
            assert len(self.min_size) == 2, "range sampling needs (min, max)"
        self.flip = is_train and inp.RANDOM_FLIP == "horizontal"
        self.device = torch.device(cfg.MODEL.DEVICE)
        self.strong_augmentation = GpuStrongAugmentation(is_train=is_train)

    def _draw_weak(self, h, w):
        if self.sample_style == "range":                       # [D2] ResizeShortestEdge.get_transform
            size = np.random.randint(self.min_size[0], self.min_size[1] + 1)
        else:
            size = np.random.choice(self.min_size)
        new_h, new_w = shortest_edge_shape(h, w, int(size), self.max_size)
        flip = bool(np.random.uniform() < 0.5) if self.flip else False      # [D2] RandomFlip(prob=0.5)
        return new_h, new_w, flip

    def __call__(self, dataset_dicts):
        weak, metas = [], []
        for d in dataset_dicts:
            img = d["image"]
            if isinstance(img, np.ndarray):
                img = torch.from_numpy(np.ascontiguousarray(img))
            img = img.to(self.device, non_blocking=True).contiguous()
            h, w = int(img.shape[0]), int(img.shape[1])
            new_h, new_w, flip = self._draw_weak(h, w)
            weak.append(resize_flip(img, new_h, new_w, flip))
            metas.append((h, w, new_h, new_w, flip))
        strong = self.strong_augmentation(weak) if self.is_train else weak
        out_q, out_k = [], []
        for d, wk, st, (h, w, new_h, new_w, flip) in zip(dataset_dicts, weak, strong, metas):
            base = {k: v for k, v in d.items() if k not in ("image", "annotations")}
            base.setdefault("height", h)
            base.setdefault("width", w)
            if self.is_train and "annotations" in d:
                annos = [a for a in d["annotations"] if a.get("iscrowd", 0) == 0]
                boxes, keep = transform_boxes([bbox_xyxy(a) for a in annos], h, w, new_h, new_w, flip)
                inst = Instances((new_h, new_w))
                inst.gt_boxes = Boxes(torch.from_numpy(boxes[keep]))
                inst.gt_classes = torch.tensor([a["category_id"] for a, k in zip(annos, keep) if k], dtype=torch.int64)
                base["instances"] = inst
            out_q.append(dict(base, image=st))
            out_k.append(dict(base, image=wk))
        return out_q, out_k


def bbox_xyxy(anno):
    """[D2] BoxMode.convert(bbox, anno["bbox_mode"], XYXY_ABS) for the two absolute modes dataset dicts carry
    (XYXY_ABS = 0: the default here; XYWH_ABS = 1: what [D2]'s COCO loader produces). Anything else is refused."""
    b = [float(v) for v in anno["bbox"]]
    mode = anno.get("bbox_mode", 0)
    mode = getattr(mode, "value", mode)
    if mode in (0, "XYXY_ABS"):
        return b
    if mode in (1, "XYWH_ABS"):
        return [b[0], b[1], b[0] + b[2], b[1] + b[3]]
    raise ValueError(f"unsupported bbox_mode {mode!r} (XYXY_ABS or XYWH_ABS expected)")
