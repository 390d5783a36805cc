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


@pytest.mark.parametrize("size,new,flip", [((67, 93), (40, 56), False), ((67, 93), (67, 50), True), ((50, 80), (100, 160), True),
                                           ((120, 75), (77, 48), False), ((200, 300), (51, 77), True), ((64, 64), (64, 64), True),
                                           ((480, 640), (800, 1067), False), ((1200, 1600), (600, 800), True)])
def test_resize_flip_matches_pillow_oracle(size, new, flip):
    from oracle import ut2_aug_oracle as A
    from ubteacher.data.dataset_mapper import resize_flip
    img = rand_img(11, *size)
    got = resize_flip(torch.from_numpy(img).cuda(), new[0], new[1], flip)
    assert np.array_equal(to_host(got), A.weak_augment(img, new[0], new[1], flip))


def test_two_crop_mapper_matches_oracle():
    """DatasetMapperTwoCropSeparate on the device == the oracle chain (weak resize + flip, box transform, strong aug of the
    weak view) under the same numpy / torch / random seeds; returns (strong dicts, weak dicts) with shared labels."""
    from oracle import ut2_aug_oracle as A
    from util_cfg import fcos_cfg
    from ubteacher.data.dataset_mapper import DatasetMapperTwoCropSeparate
    cfg = fcos_cfg(**{"INPUT.MIN_SIZE_TRAIN": (96, 160), "INPUT.MAX_SIZE_TRAIN": 220})
    mapper = DatasetMapperTwoCropSeparate(cfg, True)
    mapper.strong_augmentation.exact_noise = True
    sizes = [(120, 90), (75, 133), (200, 150), (64, 64)]
    dicts = []
    for i, (h, w) in enumerate(sizes):
        dicts.append({"image": rand_img(20 + i, h, w), "image_id": i, "file_name": f"{i}.jpg",
                      "annotations": [{"bbox": [5.0, 7.0, w - 10.0, h - 3.0], "category_id": i, "iscrowd": 0},
                                      {"bbox": [1.0, 1.0, 9.0, 9.0], "category_id": 7, "iscrowd": 1},
                                      {"bbox": [w / 2.0, h / 3.0, w / 2.0 + 11, h / 3.0 + 17], "category_id": 3, "iscrowd": 0}]})
    np.random.seed(5); torch.manual_seed(6); random.seed(7)
    q, k = mapper(dicts)
    np.random.seed(5); torch.manual_seed(6); random.seed(7)
    weak_ref, meta = [], []
    for d in dicts:
        h, w = d["image"].shape[:2]
        size = np.random.randint(96, 161)
        nh, nw = A.shortest_edge_shape(h, w, size, 220)
        flip = bool(np.random.uniform() < 0.5)
        weak_ref.append(A.weak_augment(d["image"], nh, nw, flip))
        meta.append((h, w, nh, nw, flip))
    assert any(m[4] for m in meta) and not all(m[4] for m in meta)
    for i, d in enumerate(dicts):
        h, w, nh, nw, flip = meta[i]
        assert np.array_equal(to_host(k[i]["image"]), weak_ref[i]), i
        p = A.draw_params(nh, nw)
        assert np.array_equal(to_host(q[i]["image"]), A.strong_augment(weak_ref[i], p)), i
        boxes, keep = A.transform_boxes([a["bbox"] for a in d["annotations"] if not a["iscrowd"]], h, w, nh, nw, flip)
        for view in (q[i], k[i]):
            assert view["instances"].image_size == (nh, nw) and view["height"] == h and view["width"] == w and view["image_id"] == i
            assert np.array_equal(view["instances"].gt_boxes.tensor.numpy(), boxes[keep])
            assert view["instances"].gt_classes.tolist() == [i, 3]


@pytest.mark.parametrize("size", [(70, 90), (33, 500), (300, 40), (5, 7)])
def test_gaussian_blur_fused_and_general_paths(size):
    """Box radius 0 / 1 / 2 take the fused three-pass kernels (row segments of 224 and column strips of 32 x 64 with halos
    and edge replication per pass), larger radii the pass-by-pass kernels; all equal Pillow's GaussianBlur (oracle)."""
    from oracle import ut2_aug_oracle as A
    from ubteacher.data.gpu_augmentation import GpuStrongAugmentation, box_blur_params
    img = rand_img(31, *size)
    radii = [0.1, 0.9, 2.0, 3.0, 3.4, 6.0, 9.5]
    assert sorted({box_blur_params(r)[0] for r in radii}) == [0, 1, 2, 5, 9]
    params = [{"jitter": None, "gray": False, "blur": r, "erase": []} for r in radii]
    outs = GpuStrongAugmentation()([to_dev(img) for _ in radii], params=params)
    for r, o in zip(radii, outs):
        assert np.array_equal(to_host(o), A.gaussian_blur(img, r)), (size, r)


def test_two_crop_loader_end_to_end():
    """build_detection_semisup_train_loader_two_crops with the device mapper: batches of (strong, weak) views, uint8 CHW on
    the device, strong and weak of an image share size and labels, orientation-pure batches."""
    import itertools
    from util_cfg import fcos_cfg
    from ubteacher.data.build import build_detection_semisup_train_loader_two_crops
    cfg = fcos_cfg(**{"INPUT.MIN_SIZE_TRAIN": (64, 96), "INPUT.MAX_SIZE_TRAIN": 160, "SOLVER.IMG_PER_BATCH_LABEL": 2,
                      "SOLVER.IMG_PER_BATCH_UNLABEL": 3})
    g = np.random.default_rng(4)

    def mk(n, tag, labeled):
        out = []
        for i in range(n):
            h, w = (60, 90) if g.random() < 0.5 else (90, 60)
            d = {"file_name": f"{tag}{i}", "height": h, "width": w, "image_id": i}
            if labeled:
                d["annotations"] = [{"bbox": [5.0, 6.0, w - 7.0, h - 8.0], "category_id": i % 80, "iscrowd": 0}]
            out.append(d)
        return out
    reader = lambda d: g.integers(0, 256, (d["height"], d["width"], 3), dtype=np.uint8)
    loader = build_detection_semisup_train_loader_two_crops(cfg, label_dicts=mk(12, "l", True), unlabel_dicts=mk(20, "u", False), reader=reader)
    for lq, lk, uq, uk in itertools.islice(iter(loader), 4):
        assert len(lq) == len(lk) == 2 and len(uq) == len(uk) == 3
        for q, k in zip(lq + uq, lk + uk):
            assert q["image"].is_cuda and q["image"].dtype == torch.uint8 and q["image"].shape == k["image"].shape and q["image"].shape[0] == 3
            assert q["image_id"] == k["image_id"]
        for q, k in zip(lq, lk):
            assert torch.equal(q["instances"].gt_boxes.tensor, k["instances"].gt_boxes.tensor) and len(q["instances"]) == 1
        assert len({d["width"] > d["height"] for d in lq}) == 1 and len({d["width"] > d["height"] for d in uq}) == 1
    torch.cuda.synchronize()
