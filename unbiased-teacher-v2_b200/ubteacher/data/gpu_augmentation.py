"""The reference's strong augmentation (ubteacher/data/detection_utils.py:8-46, transforms/augmentation_impl.py:7-23)
executed on the device over a whole batch of uint8 CHW images (at ~350 images/s per GPU the
PIL / CPU two-crop loader of the reference is the bottleneck).

    RandomApply([ColorJitter(0.4, 0.4, 0.4, 0.1)], p=0.8) -> RandomGrayscale(p=0.2) ->
    RandomApply([GaussianBlur([0.1, 2.0])], p=0.5) -> ToTensor -> RandomErasing x3 (value="random") -> ToPILImage

The random draws are made on the host in the order torchvision makes them (global torch RNG; Python `random` for the
blur radius), the arithmetic runs in csrc/augment.cu (Pillow's algorithms bit for bit). With ``exact_noise=True`` the
erase noise is drawn on the host like torchvision does, which keeps the RNG stream — and therefore the output —
identical to the reference pipeline for the same seeds (used by the parity tests); by default the noise is generated
on the device (same distribution) and nothing but a ~170-byte record per image crosses PCIe."""
import math
import random
import struct

import numpy as np
import torch

from .. import _C
from .._C import i64  # noqa: F401

REC = struct.Struct("<6Q2i4i4f3i2Ii12iIi")
assert REC.size == 168
ERASE_SPECS = ((0.7, (0.05, 0.2), (0.3, 3.3)), (0.5, (0.02, 0.2), (0.1, 6)), (0.3, (0.02, 0.2), (0.05, 8)))
F32 = np.float32


def box_blur_params(radius):
    """Pillow ImagingGaussianBlur -> ImagingBoxBlur parameters (libImaging/BoxBlur.c): (int radius, ww, fw) or None."""
    radius = F32(radius)
    sigma2 = F32(radius * radius / F32(3))
    L = F32(math.sqrt(12.0 * float(sigma2) + 1.0))
    l = F32(math.floor((float(L) - 1.0) / 2.0))
    a = F32((2 * l + 1) * (l * (l + 1) - 3 * sigma2))
    a = F32(float(a) / float(F32(6 * (sigma2 - (l + 1) * (l + 1)))))
    fr = F32(l + a)
    if float(fr) == 0.0:
        return None
    r = int(fr)
    ww = int(F32(16777216.0) / F32(fr * F32(2) + F32(1)))
    return r, ww, ((1 << 24) - (r * 2 + 1) * ww) // 2


class GpuStrongAugmentation:
    """Callable with the role of the reference's ``transforms.Compose`` returned by ``build_strong_augmentation``, but
    batched and device-resident: ``aug(list of uint8 [3, h, w] CUDA tensors) -> list of uint8 [3, h, w] CUDA tensors``."""

    def __init__(self, is_train=True, exact_noise=False):
        self.is_train = is_train
        self.exact_noise = exact_noise
        self.last_params = None

    # ------------------------------------------------------------------ the draws (torchvision's order)
    def draw(self, h, w):
        p = {"jitter": None, "gray": False, "blur": None, "erase": []}
        if not (0.8 < torch.rand(1)):                                   # RandomApply(p=0.8)
            order = torch.randperm(4).tolist()                          # ColorJitter.get_params
            b = float(torch.empty(1).uniform_(0.6, 1.4))
            c = float(torch.empty(1).uniform_(0.6, 1.4))
            s = float(torch.empty(1).uniform_(0.6, 1.4))
            hu = float(torch.empty(1).uniform_(-0.1, 0.1))
            p["jitter"] = (order, b, c, s, hu)
        p["gray"] = bool(torch.rand(1) < 0.2)                           # RandomGrayscale(p=0.2)
        if not (0.5 < torch.rand(1)):                                   # RandomApply([GaussianBlur], p=0.5)
            p["blur"] = random.uniform(0.1, 2.0)
        area = h * w
        for prob, scale, ratio in ERASE_SPECS:                          # RandomErasing.forward / get_params
            if not (torch.rand(1) < prob):
                continue
            log_ratio = torch.log(torch.tensor(ratio))
            for _ in range(10):
                erase_area = area * torch.empty(1).uniform_(scale[0], scale[1]).item()
                aspect = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1])).item()
                eh = int(round(math.sqrt(erase_area * aspect)))
                ew = int(round(math.sqrt(erase_area / aspect)))
                if not (eh < h and ew < w):
                    continue
                v = torch.empty([3, eh, ew], dtype=torch.float32).normal_() if self.exact_noise else None
                i = torch.randint(0, h - eh + 1, size=(1,)).item()
                j = torch.randint(0, w - ew + 1, size=(1,)).item()
                p["erase"].append((i, j, eh, ew, v))
                break
        return p

    # ------------------------------------------------------------------ the batch
    def __call__(self, images, params=None):
        if not self.is_train:
            return list(images)
        n = len(images)
        if n == 0:
            return []
        dev = images[0].device
        if dev.type != "cuda":
            raise RuntimeError("GpuStrongAugmentation runs on CUDA tensors only (there is no CPU fallback)")
        if params is None:
            params = [self.draw(int(im.shape[1]), int(im.shape[2])) for im in images]
        self.last_params = params
        outs, keep = [], []
        raw = bytearray()
        max_px = max_er = 0
        big_blur = False
        for im, p in zip(images, params):
            assert im.dtype == torch.uint8 and im.dim() == 3 and im.shape[0] == 3 and im.is_contiguous()
            h, w = int(im.shape[1]), int(im.shape[2])
            out = torch.empty_like(im)
            outs.append(out)
            order, factor, hue_shift = [-1, -1, -1, -1], [1.0, 1.0, 1.0, 0.0], 0
            if p["jitter"] is not None:
                order, b, c, s, hu = p["jitter"]
                factor = [b, c, s, hu]
                hue_shift = int(np.array(hu * 255).astype(np.int32).astype(np.uint8))
            blur = box_blur_params(p["blur"]) if p["blur"] is not None else None
            tmp = 0
            if blur is not None:
                t = torch.empty_like(im)
                keep.append(t)
                tmp = t.data_ptr()
            noise, ei, ej, eh, ew = [0, 0, 0], [0] * 3, [0] * 3, [0] * 3, [0] * 3
            for k, (i, j, hh, ww_, v) in enumerate(p["erase"]):
                ei[k], ej[k], eh[k], ew[k] = i, j, hh, ww_
                if v is not None:
                    vd = torch.as_tensor(v, dtype=torch.float32).contiguous().to(dev, non_blocking=True)
                    keep.append(vd)
                    noise[k] = vd.data_ptr()
            seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if not self.exact_noise else 0
            br, bww, bfw = blur if blur is not None else (-1, 0, 0)
            big_blur = big_blur or br > 2
            raw += REC.pack(im.data_ptr(), out.data_ptr(), tmp, noise[0], noise[1], noise[2], h, w, *order, *factor, hue_shift,
                            int(p["gray"]), br, bww, bfw, len(p["erase"]), *ei, *ej, *eh, *ew, seed, 0)
            max_px = max(max_px, h * w)
            max_er = max(max_er, len(p["erase"]))
        table = torch.frombuffer(raw, dtype=torch.uint8).clone().pin_memory().to(dev, non_blocking=True)
        ws = torch.empty((n, 4), dtype=torch.int64, device=dev)
        _C.counted_call("ut2_strong_augment_u8", table, n, max_px, max_er + (256 if big_blur else 0), ws)
        _C.launch_count += 12 + max_er + (6 if big_blur else 0)
        self._keep = (keep, table, ws)      # alive until the next call (stream-ordered use)
        return outs
