"""Pins oracle/ut2_aug_oracle.py against the real thing: Pillow + torchvision are installed here, so the reference's own
strong augmentation (ubteacher/data/detection_utils.py:8-46) is executed on the CPU and the restatement must agree
bit for bit, op by op and end to end (same torch / random seeds => same draws)."""
import os
import random
import sys

import numpy as np
import pytest
import torch

PIL = pytest.importorskip("PIL")
tvt = pytest.importorskip("torchvision.transforms")
from PIL import Image, ImageFilter  # noqa: E402

from oracle import ut2_aug_oracle as A  # noqa: E402

REF = "/root/reference"


def rand_img(seed, h=67, w=93):
    g = np.random.default_rng(seed)
    base = g.integers(0, 256, (h, w, 3), dtype=np.uint8)
    # smooth half of it so that blur / hue see more than white noise
    base[:, : w // 2] = (np.linspace(0, 255, h)[:, None, None] * np.ones((1, w // 2, 3)) * g.random(3)).astype(np.uint8)
    return base


def test_hsv_roundtrip_all_colours():
    v = np.arange(1 << 24, dtype=np.uint32)
    rgb = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], axis=-1).astype(np.uint8).reshape(4096, 4096, 3)
    hsv = np.array(Image.fromarray(rgb, "RGB").convert("HSV"))
    assert np.array_equal(A.rgb2hsv(rgb), hsv)
    back = np.array(Image.fromarray(hsv, "HSV").convert("RGB"))
    assert np.array_equal(A.hsv2rgb(hsv), back)


def test_hsv_to_rgb_all_triples_sampled():
    g = np.random.default_rng(0)
    hsv = g.integers(0, 256, (2048, 2048, 3), dtype=np.uint8)
    assert np.array_equal(A.hsv2rgb(hsv), np.array(Image.fromarray(hsv, "HSV").convert("RGB")))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_colour_ops_match_torchvision_pil(seed):
    import torchvision.transforms.functional as F
    img = rand_img(seed)
    pil = Image.fromarray(img, "RGB")
    assert np.array_equal(A.to_L(img), np.array(pil.convert("L")))
    for f in (0.6, 0.83, 1.0, 1.17, 1.4):
        assert np.array_equal(A.adjust_brightness(img, f), np.array(F.adjust_brightness(pil, f))), f
        assert np.array_equal(A.adjust_contrast(img, f), np.array(F.adjust_contrast(pil, f))), f
        assert np.array_equal(A.adjust_saturation(img, f), np.array(F.adjust_saturation(pil, f))), f
    for f in (-0.1, -0.031, 0.0, 0.004, 0.07, 0.1):
        assert np.array_equal(A.adjust_hue(img, f), np.array(F.adjust_hue(pil, f))), f
    assert np.array_equal(A.to_grayscale3(img), np.array(F.rgb_to_grayscale(pil, 3)))


@pytest.mark.parametrize("size", [(67, 93), (5, 200), (40, 3)])
def test_gaussian_blur_matches_pillow(size):
    img = rand_img(5, *size)
    pil = Image.fromarray(img, "RGB")
    for radius in (0.1, 0.2, 0.35, 0.5, 0.77, 1.0, 1.3, 1.61, 1.99, 2.0):
        want = np.array(pil.filter(ImageFilter.GaussianBlur(radius=radius)))
        assert np.array_equal(A.gaussian_blur(img, radius), want), radius


def test_totensor_topil_and_byte_cast():
    x = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    t = tvt.ToTensor()(Image.fromarray(x, "RGB"))
    back = np.array(tvt.ToPILImage()(t))
    assert np.array_equal(A.totensor_erase_topil(x, []), back)
    f = torch.tensor([-700.3, -256.0, -255.9, -1.5, -0.4, 0.0, 0.99, 1.0, 254.99998, 255.0, 256.0, 300.7, 511.9, 1000.2])
    assert np.array_equal(A.float_to_byte(f.numpy()), f.byte().numpy())


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("seed", list(range(12)))
def test_end_to_end_matches_reference_pipeline(seed):
    import importlib.util
    import types

    def load(rel, name):       # execute the reference FILES under private module names (its package imports Detectron2)
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    saved = {k: sys.modules.get(k) for k in ("ubteacher", "ubteacher.data", "ubteacher.data.transforms",
                                             "ubteacher.data.transforms.augmentation_impl")}
    try:
        impl = load("ubteacher/data/transforms/augmentation_impl.py", "_ref_augmentation_impl")
        for k in list(saved)[:3]:
            sys.modules[k] = types.ModuleType(k)
        sys.modules["ubteacher.data.transforms.augmentation_impl"] = impl
        build_strong_augmentation = load("ubteacher/data/detection_utils.py", "_ref_detection_utils").build_strong_augmentation
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    aug = build_strong_augmentation(None, True)
    img = rand_img(100 + seed, 61 + 3 * seed, 80 + 5 * seed)
    torch.manual_seed(seed); random.seed(seed)
    want = np.array(aug(Image.fromarray(img, "RGB")))
    torch.manual_seed(seed); random.seed(seed)
    p = A.draw_params(img.shape[0], img.shape[1])
    got = A.strong_augment(img, p)
    assert np.array_equal(got, want), (seed, p["jitter"], p["gray"], p["blur"], [e[:4] for e in p["erase"]])


@pytest.mark.parametrize("size,new", [((67, 93), (40, 56)), ((67, 93), (67, 50)), ((50, 80), (100, 160)), ((120, 75), (77, 48)),
                                      ((33, 41), (90, 41)), ((200, 300), (51, 77)), ((64, 64), (63, 65))])
def test_pil_bilinear_resize_bit_exact(size, new):
    img = rand_img(9, *size)
    want = np.array(Image.fromarray(img, "RGB").resize((new[1], new[0]), Image.BILINEAR))
    assert np.array_equal(A.pil_resize_bilinear(img, new[1], new[0]), want)


def test_shortest_edge_and_box_transform():
    assert A.shortest_edge_shape(480, 640, 800, 1333) == (800, 1067)
    assert A.shortest_edge_shape(427, 640, 800, 1333) == (800, 1199)
    assert A.shortest_edge_shape(300, 1000, 800, 1333) == (400, 1333)
    assert A.shortest_edge_shape(640, 480, 600, 1333) == (800, 600)
    b, keep = A.transform_boxes([[10, 20, 110, 220], [600, 0, 700, 50], [5, 5, 5, 9]], 480, 640, 240, 320, True)
    assert np.allclose(b[0], [320 - 55, 10, 320 - 5, 110]) and np.allclose(b[1], [0, 0, 20, 25]) and list(keep) == [True, True, False]
