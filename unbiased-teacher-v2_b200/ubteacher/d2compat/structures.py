"""Minimal stand-ins for the Detectron2 v0.6 containers the UT2 hot path touches
(``Boxes``, ``Instances``, ``ImageList``, ``cat``, ``ShapeSpec``).

Detectron2 is an un-vendored dependency of the reference (reference README.md:27); when the real
package is importable these names are re-exported from it, otherwise the shims below are used.
Only the behaviour the reference call sites rely on is provided.
"""
import itertools
from collections import namedtuple
from typing import Any, Dict, List, Tuple, Union

import torch

try:  # pragma: no cover - exercised only where detectron2 exists
    from detectron2.layers import ShapeSpec, cat  # noqa: F401
    from detectron2.structures import Boxes, ImageList, Instances  # noqa: F401

    HAVE_D2 = True
except Exception:  # noqa: BLE001
    HAVE_D2 = False


if not HAVE_D2:

    class ShapeSpec(namedtuple("_ShapeSpec", ["channels", "height", "width", "stride"])):
        def __new__(cls, channels=None, height=None, width=None, stride=None):
            return super().__new__(cls, channels, height, width, stride)

    def cat(tensors: List[torch.Tensor], dim: int = 0):
        assert isinstance(tensors, (list, tuple))
        if len(tensors) == 1:
            return tensors[0]
        return torch.cat(tensors, dim)

    class Boxes:
        """[N, 4] xyxy absolute boxes."""

        def __init__(self, tensor: torch.Tensor):
            if not isinstance(tensor, torch.Tensor):
                tensor = torch.as_tensor(tensor, dtype=torch.float32)
            else:
                tensor = tensor.to(torch.float32)
            if tensor.numel() == 0:
                tensor = tensor.reshape((-1, 4)).to(dtype=torch.float32)
            assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
            self.tensor = tensor

        def clone(self):
            return Boxes(self.tensor.clone())

        def to(self, *args, **kwargs):
            return Boxes(self.tensor.to(*args, **kwargs))

        def area(self):
            box = self.tensor
            return (box[:, 2] - box[:, 0]) * (box[:, 3] - box[:, 1])

        def clip(self, box_size: Tuple[int, int]):
            h, w = box_size
            x1 = self.tensor[:, 0].clamp(min=0, max=w)
            y1 = self.tensor[:, 1].clamp(min=0, max=h)
            x2 = self.tensor[:, 2].clamp(min=0, max=w)
            y2 = self.tensor[:, 3].clamp(min=0, max=h)
            self.tensor = torch.stack((x1, y1, x2, y2), dim=-1)

        def scale(self, scale_x: float, scale_y: float):
            self.tensor = self.tensor.clone()
            self.tensor[:, 0::2] *= scale_x
            self.tensor[:, 1::2] *= scale_y

        def nonempty(self, threshold: float = 0.0):
            box = self.tensor
            return ((box[:, 2] - box[:, 0]) > threshold) & ((box[:, 3] - box[:, 1]) > threshold)

        def __getitem__(self, item):
            if isinstance(item, int):
                return Boxes(self.tensor[item].view(1, -1))
            b = self.tensor[item]
            assert b.dim() == 2, "Indexing on Boxes with {} failed".format(item)
            return Boxes(b)

        def __len__(self):
            return self.tensor.shape[0]

        def __repr__(self):
            return "Boxes(" + str(self.tensor) + ")"

        @classmethod
        def cat(cls, boxes_list):
            assert isinstance(boxes_list, (list, tuple))
            if len(boxes_list) == 0:
                return cls(torch.empty(0))
            return cls(torch.cat([b.tensor for b in boxes_list], dim=0))

        @property
        def device(self):
            return self.tensor.device

        def __iter__(self):
            yield from self.tensor

    class Instances:
        """Per-image field container; every field has the same length."""

        def __init__(self, image_size: Tuple[int, int], **kwargs: Any):
            self._image_size = image_size
            self._fields: Dict[str, Any] = {}
            for k, v in kwargs.items():
                self.set(k, v)

        @property
        def image_size(self):
            return self._image_size

        def __setattr__(self, name, val):
            if name.startswith("_"):
                super().__setattr__(name, val)
            else:
                self.set(name, val)

        def __getattr__(self, name):
            if name == "_fields" or name not in self._fields:
                raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
            return self._fields[name]

        def set(self, name, value):
            data_len = len(value)
            if len(self._fields):
                assert len(self) == data_len, "Adding a field of length {} to a Instances of length {}".format(
                    data_len, len(self)
                )
            self._fields[name] = value

        def has(self, name):
            return name in self._fields

        def remove(self, name):
            del self._fields[name]

        def get(self, name):
            return self._fields[name]

        def get_fields(self):
            return self._fields

        def to(self, *args, **kwargs):
            ret = Instances(self._image_size)
            for k, v in self._fields.items():
                if hasattr(v, "to"):
                    v = v.to(*args, **kwargs)
                ret.set(k, v)
            return ret

        def __getitem__(self, item):
            if type(item) == int:
                if item >= len(self) or item < -len(self):
                    raise IndexError("Instances index out of range!")
                item = slice(item, None, len(self))
            ret = Instances(self._image_size)
            for k, v in self._fields.items():
                ret.set(k, v[item])
            return ret

        def __len__(self):
            for v in self._fields.values():
                return v.__len__()
            raise NotImplementedError("Empty Instances does not support __len__!")

        def __iter__(self):
            raise NotImplementedError("`Instances` object is not iterable!")

        @staticmethod
        def cat(instance_lists):
            assert all(isinstance(i, Instances) for i in instance_lists)
            assert len(instance_lists) > 0
            if len(instance_lists) == 1:
                return instance_lists[0]
            image_size = instance_lists[0].image_size
            ret = Instances(image_size)
            for k in instance_lists[0]._fields.keys():
                values = [i.get(k) for i in instance_lists]
                v0 = values[0]
                if isinstance(v0, torch.Tensor):
                    values = torch.cat(values, dim=0)
                elif isinstance(v0, list):
                    values = list(itertools.chain(*values))
                elif hasattr(type(v0), "cat"):
                    values = type(v0).cat(values)
                else:
                    raise ValueError("Unsupported type {} for concatenation".format(type(v0)))
                ret.set(k, values)
            return ret

        def __repr__(self):
            s = self.__class__.__name__ + "("
            s += "num_instances={}, ".format(len(self) if len(self._fields) else 0)
            s += "image_height={}, image_width={}, ".format(self._image_size[0], self._image_size[1])
            s += "fields=[{}])".format(", ".join(f"{k}: {v}" for k, v in self._fields.items()))
            return s

    class ImageList:
        """Batched, zero-padded images + the un-padded (h, w) of each."""

        def __init__(self, tensor: torch.Tensor, image_sizes: List[Tuple[int, int]]):
            self.tensor = tensor
            self.image_sizes = image_sizes

        def __len__(self):
            return len(self.image_sizes)

        def __getitem__(self, idx):
            size = self.image_sizes[idx]
            return self.tensor[idx, ..., : size[0], : size[1]]

        def to(self, *args, **kwargs):
            return ImageList(self.tensor.to(*args, **kwargs), self.image_sizes)

        @property
        def device(self):
            return self.tensor.device

        @staticmethod
        def from_tensors(tensors, size_divisibility: int = 0, pad_value: float = 0.0):
            assert len(tensors) > 0
            image_sizes = [(im.shape[-2], im.shape[-1]) for im in tensors]
            max_h = max(s[0] for s in image_sizes)
            max_w = max(s[1] for s in image_sizes)
            if size_divisibility > 1:
                d = size_divisibility
                max_h = (max_h + d - 1) // d * d
                max_w = (max_w + d - 1) // d * d
            batch_shape = [len(tensors)] + list(tensors[0].shape[:-2]) + [max_h, max_w]
            batched = tensors[0].new_full(batch_shape, pad_value)
            for img, pad_img in zip(tensors, batched):
                pad_img[..., : img.shape[-2], : img.shape[-1]].copy_(img)
            return ImageList(batched.contiguous(), image_sizes)


def detector_postprocess(results, output_height, output_width):
    """[D2] v0.6 ``modeling.postprocessing.detector_postprocess`` for box-only results (the part the reference's
    wrapper ``ubteacher/modeling/one_stage_detector.py:16-43`` delegates to; UT2 has no masks / keypoints / beziers):
    rescale ``pred_boxes`` (or ``proposal_boxes``) from the network input resolution to (output_height, output_width),
    clip, and drop empty boxes."""
    scale_x = float(output_width) / results.image_size[1]
    scale_y = float(output_height) / results.image_size[0]
    out = Instances((int(output_height), int(output_width)), **results.get_fields())
    if out.has("pred_boxes"):
        boxes = out.pred_boxes = out.pred_boxes.clone()       # ([D2] rescales the caller's Boxes in place; a copy here)
    elif out.has("proposal_boxes"):
        boxes = out.proposal_boxes = out.proposal_boxes.clone()
    else:
        return out
    boxes.scale(scale_x, scale_y)
    boxes.clip(out.image_size)
    return out[boxes.nonempty()]
