"""GPU parity of the device strong augmentation (csrc/augment.cu through the C ABI) against the CPU oracle
(oracle/ut2_aug_oracle.py, itself pinned bit-exactly against the reference pipeline in tests/test_oracle_aug.py):
uint8 images, so everything is bit-exact."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rand_img(seed, h, w):
    g = np.random.default_rng(seed)
    base = g.integers(0, 256, (h, w, 3), dtype=np.uint8)
    base[:, : w // 2] = (np.linspace(0, 255, h)[:, None, None] * np.ones((1, w // 2, 3)) * g.random(3)).astype(np.uint8)
    return base


def to_dev(img):
    return torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1))).cuda()


def to_host(t):
    return t.cpu().numpy().transpose(1, 2, 0)


def test_batch_matches_oracle_bit_exact_same_seeds():
    """Same torch / random seeds => the device pipeline draws what the reference pipeline draws (exact_noise mode) and
    produces the oracle's image bit for bit; ragged sizes, one launch sequence for the whole batch."""
    from oracle import ut2_aug_oracle as A
    from ubteacher.data.gpu_augmentation import GpuStrongAugmentation
    sizes = [(61, 80), (97, 131), (40, 203), (128, 128), (75, 33), (64, 300), (211, 157), (50, 50)] * 3
    imgs = [rand_img(i, h, w) for i, (h, w) in enumerate(sizes)]
    torch.manual_seed(1234); random.seed(99)
    want = []
    for im in imgs:
        p = A.draw_params(im.shape[0], im.shape[1])
        want.append((A.strong_augment(im, p), p))
    torch.manual_seed(1234); random.seed(99)
    aug = GpuStrongAugmentation(exact_noise=True)
    got = aug([to_dev(im) for im in imgs])
    kinds = set()
    for i, (g, (w_, p)) in enumerate(zip(got, want)):
        assert np.array_equal(to_host(g), w_), (i, sizes[i], p["jitter"], p["gray"], p["blur"], [e[:4] for e in p["erase"]])
        kinds |= {"jitter"} if p["jitter"] else set()
        kinds |= {"gray"} if p["gray"] else set()
        kinds |= {"blur"} if p["blur"] else set()
        kinds |= {"erase%d" % len(p["erase"])}
    assert {"jitter", "gray", "blur"} <= kinds and len([k for k in kinds if k.startswith("erase")]) >= 2


@pytest.mark.parametrize("order", [[0, 1, 2, 3], [3, 2, 1, 0], [1, 3, 0, 2], [2, 0, 3, 1]])
def test_every_op_order_and_extreme_factors(order):
    from oracle import ut2_aug_oracle as A
    from ubteacher.data.gpu_augmentation import GpuStrongAugmentation
    g = np.random.default_rng(7)
    imgs, params = [], []
    for k, (b, c, s, hu) in enumerate([(0.6, 0.6, 0.6, -0.1), (1.4, 1.4, 1.4, 0.1), (1.0, 1.0, 1.0, 0.0), (0.61, 1.39, 0.99, 0.0039)]):
        imgs.append(g.integers(0, 256, (70 + k, 90 + 3 * k, 3), dtype=np.uint8))
        params.append({"jitter": (order, b, c, s, hu), "gray": k == 1, "blur": [None, 0.1, 2.0, 1.234][k], "erase": []})
    got = GpuStrongAugmentation()([to_dev(im) for im in imgs], params=params)
    for im, p, o in zip(imgs, params, got):
        assert np.array_equal(to_host(o), A.strong_augment(im, p)), p


def test_hue_kernel_over_all_colours():
    """4096 x 4096 image holding every 24-bit colour once: the HSV round trip of the device equals Pillow's (via the oracle)."""
    from oracle import ut2_aug_oracle as A
    from ubteacher.data.gpu_augmentation import GpuStrongAugmentation
    v = np.arange(1 << 24, dtype=np.uint32)
    rgb = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], axis=-1).astype(np.uint8).reshape(4096, 4096, 3)
    for hu in (0.0, 0.0731, -0.1):
        p = {"jitter": ([3, 0, 1, 2], 1.0, 1.0, 1.0, hu), "gray": False, "blur": None, "erase": []}
        p2 = dict(p, jitter=([3], 1.0, 1.0, 1.0, hu))
        got = GpuStrongAugmentation()([to_dev(rgb)], params=[dict(p, jitter=([3, -1, -1, -1], 1.0, 1.0, 1.0, hu))])[0]
        assert np.array_equal(to_host(got), A.strong_augment(rgb, p2)), hu


def test_device_noise_statistics_and_untouched_pixels():
    from ubteacher.data.gpu_augmentation import GpuStrongAugmentation
    img = rand_img(3, 300, 400)
    p = {"jitter": None, "gray": False, "blur": None, "erase": [(10, 20, 120, 200, None), (100, 150, 90, 60, None)]}
    out = to_host(GpuStrongAugmentation()([to_dev(img)], params=[p])[0])
    mask = np.zeros(img.shape[:2], bool)
    mask[10:130, 20:220] = True
    mask[100:190, 150:210] = True
    assert np.array_equal(out[~mask], img[~mask])
    inside = out[mask].astype(np.float64)
    # (255 * N(0,1)) mod 256 is close to uniform over the 256 byte values
    hist = np.bincount(out[mask].ravel(), minlength=256) / inside.size
    assert hist.min() > 0.5 / 256 and hist.max() < 1.6 / 256 and abs(inside.mean() - 127.5) < 2.0
    # the two regions use different noise, channels too
    assert not np.array_equal(out[10:100, 20:80, 0], out[10:100, 20:80, 1])


def test_full_size_batch_runs_and_is_deterministic():
    from ubteacher.data.gpu_augmentation import GpuStrongAugmentation
    g = torch.Generator().manual_seed(0)
    imgs = [torch.randint(0, 256, (3, 800, 1333), generator=g, dtype=torch.uint8).cuda() for _ in range(4)]
    aug = GpuStrongAugmentation()
    torch.manual_seed(5); random.seed(5)
    a = aug(imgs)
    pa = aug.last_params
    b = aug(imgs, params=pa)
    torch.cuda.synchronize()
    assert all(x.shape == (3, 800, 1333) and x.dtype == torch.uint8 for x in a)
    # same parameters (the hashed-noise seed is redrawn, so compare outside the erased regions only)
    for x, y, p in zip(a, b, pa):
        m = torch.ones(800, 1333, dtype=torch.bool, device="cuda")
        for (i, j, h, w, _) in p["erase"]:
            m[i:i + h, j:j + w] = False
        assert torch.equal(x[:, m], y[:, m])
