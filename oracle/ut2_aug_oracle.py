"""TEST INFRASTRUCTURE (CPU oracle) — restatement of the reference's strong augmentation
(ubteacher/data/detection_utils.py:8-46, ubteacher/data/transforms/augmentation_impl.py:7-23), i.e. of the Pillow /
torchvision algorithms it calls, in numpy integer / float32 arithmetic:

    RandomApply([ColorJitter(0.4, 0.4, 0.4, 0.1)], p=0.8) -> RandomGrayscale(p=0.2) ->
    RandomApply([GaussianBlur([0.1, 2.0])], p=0.5) -> ToTensor -> RandomErasing x3 (value="random") -> ToPILImage

Third-party arithmetic restated here (not under /root/reference): Pillow 12.2.0 (ImageEnhance -> Image.blend,
Image.convert "L"/"HSV"/"RGB", ImageStat mean, ImageFilter.GaussianBlur = 3-pass extended box blur) and torchvision
0.26 (ColorJitter / RandomGrayscale / RandomErasing parameter draws and their order, ToTensor, ToPILImage). Pinned: Pillow
and torchvision are installed in the build container, so tests/test_oracle_aug.py runs the reference's own
build_strong_augmentation on seeded images and requires bit-identical output, and checks the HSV round trip over all
2^24 colours. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.

Images are uint8 arrays [H, W, 3] in the channel order the reference hands to Pillow (it wraps its BGR array as "RGB",
dataset_mapper.py:116 — channel 0 plays Pillow's R)."""
import math
import random

import numpy as np
import torch

F32 = np.float32


# ------------------------------------------------------------------------------------------------ Pillow pieces
def to_L(img):
    """Image.convert("L"): ITU-R 601-2 luma, fixed point (libImaging/Convert.c rgb2l)."""
    r, g, b = (img[..., i].astype(np.uint32) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(deg, img, alpha):
    """Image.blend(deg, img, alpha) (libImaging/Blend.c): float32 `in1 + alpha * (in2 - in1)`, truncated; clipped only
    when extrapolating."""
    a = F32(alpha)
    t = deg.astype(F32) + a * (img.astype(np.int32) - deg.astype(np.int32)).astype(F32)
    if 0.0 <= alpha <= 1.0:
        return t.astype(np.int32).astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t.astype(np.int32))).astype(np.uint8)


def adjust_brightness(img, f):
    return blend(np.zeros_like(img), img, f)


def adjust_contrast(img, f):
    L = to_L(img)
    hist = np.bincount(L.ravel(), minlength=256).astype(np.int64)
    mean = int(float((hist * np.arange(256)).sum()) / float(L.size) + 0.5)        # ImageStat.Stat(...).mean
    return blend(np.full_like(img, mean), img, f)


def adjust_saturation(img, f):
    return blend(np.repeat(to_L(img)[..., None], 3, axis=2), img, f)


def rgb2hsv(img):
    """libImaging/Convert.c rgb2hsv_row (float arithmetic, truncating casts)."""
    r, g, b = (img[..., i] for i in range(3))
    maxc = np.maximum(r, np.maximum(g, b))
    minc = np.minimum(r, np.minimum(g, b))
    cr = (maxc.astype(np.int32) - minc.astype(np.int32)).astype(F32)
    safe = np.where(cr == 0, F32(1), cr)
    s = cr / np.where(maxc == 0, F32(1), maxc.astype(F32))
    rc = (maxc.astype(np.int32) - r).astype(F32) / safe
    gc = (maxc.astype(np.int32) - g).astype(F32) / safe
    bc = (maxc.astype(np.int32) - b).astype(F32) / safe
    rd, gd, bd = rc.astype(np.float64), gc.astype(np.float64), bc.astype(np.float64)
    # `h = 2.0 + rc - bc`: double literals promote the sum to double, the assignment rounds it to `float h`
    h = np.where(r == maxc, (bc - gc).astype(np.float64), np.where(g == maxc, 2.0 + rd - bd, 4.0 + gd - rd)).astype(F32)
    h = np.fmod(h.astype(np.float64) / 6.0 + 1.0, 1.0).astype(F32)      # `float h`: rounded to single precision here
    uh = np.clip((h.astype(np.float64) * 255.0).astype(np.int32), 0, 255)
    us = np.clip((s.astype(np.float64) * 255.0).astype(np.int32), 0, 255)
    gray = minc == maxc
    return np.stack([np.where(gray, 0, uh), np.where(gray, 0, us), maxc], axis=-1).astype(np.uint8)


def hsv2rgb(hsv):
    """libImaging/Convert.c hsv2rgb_row."""
    h, s, v = (hsv[..., i] for i in range(3))
    hf = h.astype(np.float64) * 6.0 / 255.0
    i = np.floor(hf)
    f = (hf - i).astype(F32).astype(np.float64)
    fs = (s.astype(F32) / F32(255.0)).astype(np.float64)
    vf = v.astype(np.float64)
    rnd = lambda x: np.clip(np.round(x), 0, 255).astype(np.uint8)        # noqa: E731  (C round: half away from zero)
    p = rnd(vf * (1.0 - fs))
    q = rnd(vf * (1.0 - fs * f))
    t = rnd(vf * (1.0 - fs * (1.0 - f)))
    k = i.astype(np.int32) % 6
    r = np.choose(k, [v, q, p, p, t, v])
    g = np.choose(k, [t, v, v, q, p, p])
    b = np.choose(k, [p, p, t, v, v, q])
    gray = s == 0
    return np.stack([np.where(gray, v, r), np.where(gray, v, g), np.where(gray, v, b)], axis=-1).astype(np.uint8)


def adjust_hue(img, f):
    """torchvision _functional_pil.adjust_hue: uint8 wrap-around shift of Pillow's H channel."""
    hsv = rgb2hsv(img)
    shift = np.array(f * 255).astype(np.int32).astype(np.uint8)
    hsv[..., 0] = hsv[..., 0] + shift            # uint8 wrap-around
    return hsv2rgb(hsv)


def to_grayscale3(img):
    return np.repeat(to_L(img)[..., None], 3, axis=2)


def gaussian_box_radius(radius, passes=3):
    """libImaging/BoxBlur.c _gaussian_blur_radius (float variables, double literals)."""
    radius = F32(radius)
    sigma2 = F32(radius * radius / F32(passes))
    L = F32(math.sqrt(12.0 * float(sigma2) + 1.0))
    l = F32(math.floor((float(L) - 1.0) / 2.0))
    a = F32((2 * l + 1) * (l * (l + 1) - 3 * sigma2))
    a = F32(float(a) / float(F32(6 * (sigma2 - (l + 1) * (l + 1)))))
    return F32(l + a)


def box_weights(float_radius):
    r = int(float_radius)
    ww = int(F32(16777216.0) / F32(F32(float_radius) * F32(2) + F32(1)))
    fw = ((1 << 24) - (r * 2 + 1) * ww) // 2
    return r, ww, fw


def box_blur_h(img, float_radius):
    """One ImagingHorizontalBoxBlur pass: out = (ww * sum_{|k|<=r} in[x+k] + fw * (in[x-r-1] + in[x+r+1]) + 2^23) >> 24,
    edge pixels replicated."""
    r, ww, fw = box_weights(float_radius)
    H, W, _ = img.shape
    xs = np.arange(W)
    acc = np.zeros(img.shape, dtype=np.uint64)
    for k in range(-r, r + 1):
        acc += img[:, np.clip(xs + k, 0, W - 1)].astype(np.uint64)
    far = img[:, np.clip(xs - r - 1, 0, W - 1)].astype(np.uint64) + img[:, np.clip(xs + r + 1, 0, W - 1)].astype(np.uint64)
    return ((acc * ww + far * fw + (1 << 23)) >> 24).astype(np.uint8)


def gaussian_blur(img, radius):
    """ImageFilter.GaussianBlur(radius): three horizontal box passes, transpose, three more, transpose back."""
    fr = gaussian_box_radius(radius)
    if float(fr) == 0.0:
        return img.copy()
    out = img
    for _ in range(3):
        out = box_blur_h(out, fr)
    out = out.transpose(1, 0, 2)
    for _ in range(3):
        out = box_blur_h(out, fr)
    return np.ascontiguousarray(out.transpose(1, 0, 2))


# ------------------------------------------------------------------------------------------------ torchvision pieces
def float_to_byte(x):
    """Tensor.byte() of a float32 CPU tensor: truncate toward zero, keep the low 8 bits."""
    return (np.trunc(x.astype(np.float64)).astype(np.int64) & 0xFF).astype(np.uint8)


def totensor_erase_topil(img, erases):
    """ToTensor -> RandomErasing(value="random") x n -> ToPILImage. erases: [(i, j, h, w, noise float32 [3, h, w])]."""
    t = img.astype(F32) / F32(255)
    for (i, j, h, w, v) in erases:
        t[i:i + h, j:j + w, :] = np.asarray(v, dtype=F32).transpose(1, 2, 0)
    return float_to_byte(t * F32(255))


ERASE_SPECS = ((0.7, (0.05, 0.2), (0.3, 3.3)), (0.5, (0.02, 0.2), (0.1, 6)), (0.3, (0.02, 0.2), (0.05, 8)))


def draw_params(h, w):
    """The random draws of one call of the reference pipeline, in the reference's order, from the global torch RNG
    (torchvision) and Python's `random` (GaussianBlur). Returns a dict of plain numbers (+ erase noise tensors)."""
    p = {"jitter": None, "gray": False, "blur": None, "erase": []}
    if not (0.8 < torch.rand(1)):
        order = torch.randperm(4).tolist()
        b = float(torch.empty(1).uniform_(0.6, 1.4))
        c = float(torch.empty(1).uniform_(0.6, 1.4))
        s = float(torch.empty(1).uniform_(0.6, 1.4))
        hu = float(torch.empty(1).uniform_(-0.1, 0.1))
        p["jitter"] = (order, b, c, s, hu)
    p["gray"] = bool(torch.rand(1) < 0.2)
    if not (0.5 < torch.rand(1)):
        p["blur"] = random.uniform(0.1, 2.0)
    area = h * w
    for prob, scale, ratio in ERASE_SPECS:
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
            v = torch.empty([3, eh, ew], dtype=torch.float32).normal_()
            i = torch.randint(0, h - eh + 1, size=(1,)).item()
            j = torch.randint(0, w - ew + 1, size=(1,)).item()
            p["erase"].append((i, j, eh, ew, v.numpy()))
            break
    return p


def strong_augment(img, p):
    out = img
    if p["jitter"] is not None:
        order, b, c, s, hu = p["jitter"]
        for fn in order:
            out = (adjust_brightness, adjust_contrast, adjust_saturation, adjust_hue)[fn](out, (b, c, s, hu)[fn])
    if p["gray"]:
        out = to_grayscale3(out)
    if p["blur"] is not None:
        out = gaussian_blur(out, p["blur"])
    return totensor_erase_topil(out, p["erase"])


# ------------------------------------------------------------------------------------------------ weak augmentation
# [D2] ResizeShortestEdge + RandomFlip (dataset_mapper.py:88-91 through detectron2.data.transforms; Detectron2 is not on
# disk: the parameter logic below restates v0.6 `ResizeShortestEdge.get_output_shape`, `ResizeTransform`, `HFlipTransform`
# and `transform_instance_annotations` — unpinned) on top of Pillow's `Image.resize(..., BILINEAR)` (libImaging/Resample.c:
# antialiased separable convolution with 22-bit fixed-point coefficients, uint8 between the passes — pinned against Pillow).
PRECISION_BITS = 32 - 8 - 2


def resample_coeffs(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (support 1.0) over the whole input axis."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        x = np.arange(xmax)
        w = np.maximum(1.0 - np.abs((x + xmin - center + 0.5) * ss), 0.0)
        ww = 0.0
        for v in w:                       # the C loop adds left to right (numpy's sum would go pairwise from 8 terms on)
            ww += float(v)
        if ww != 0.0:
            w = w / ww
        kk[xx, :xmax] = np.where(w < 0, (-0.5 + w * (1 << PRECISION_BITS)).astype(np.int64), (0.5 + w * (1 << PRECISION_BITS)).astype(np.int64))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis1(img, out_size):
    """One pass along axis 1 of a uint8 [A, in, C] array."""
    bounds, kk = resample_coeffs(img.shape[1], out_size)
    out = np.empty((img.shape[0], out_size, img.shape[2]), dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, xmax = bounds[xx]
        acc = (1 << (PRECISION_BITS - 1)) + (src[:, xmin:xmin + xmax, :] * kk[xx, :xmax, None]).sum(axis=1)
        out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return out


def pil_resize_bilinear(img, new_w, new_h):
    """Image.resize((new_w, new_h), BILINEAR) for uint8 [H, W, 3]: horizontal pass, then vertical pass."""
    out = img
    if new_w != img.shape[1]:
        out = _resample_axis1(out, new_w)
    if new_h != img.shape[0]:
        out = _resample_axis1(out.transpose(1, 0, 2), new_h).transpose(1, 0, 2)
    return np.ascontiguousarray(out)


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


def weak_augment(img, new_h, new_w, flip):
    out = pil_resize_bilinear(img, new_w, new_h) if (new_h, new_w) != img.shape[:2] else img
    return np.ascontiguousarray(out[:, ::-1]) if flip else out


def transform_boxes(boxes, h, w, new_h, new_w, flip):
    """[D2] transform_instance_annotations for XYXY_ABS boxes: ResizeTransform.apply_coords, HFlipTransform.apply_coords on
    the four corners, min / max, clip to the image; then filter_empty_instances (side > 1e-5). Returns (boxes, keep)."""
    b = np.asarray(boxes, dtype=np.float64).reshape(-1, 4).copy()
    sx, sy = new_w * 1.0 / w, new_h * 1.0 / h
    x0, y0, x1, y1 = b[:, 0] * sx, b[:, 1] * sy, b[:, 2] * sx, b[:, 3] * sy
    if flip:
        x0, x1 = new_w - x1, new_w - x0
    out = np.stack([np.minimum(x0, x1), np.minimum(y0, y1), np.maximum(x0, x1), np.maximum(y0, y1)], axis=1).clip(min=0)
    out = np.minimum(out, np.array([new_w, new_h, new_w, new_h], dtype=np.float64))
    keep = ((out[:, 2] - out[:, 0]) > 1e-5) & ((out[:, 3] - out[:, 1]) > 1e-5)
    return out.astype(np.float32), keep
