"""Throughput of the two-crop strong augmentation: device kernels (csrc/augment.cu) vs the reference's PIL / torchvision
pipeline on the host cores (run under gpurun). Images: uint8 3x800x1333, the reference's draws (p = .8 / .2 / .5 / .7 / .5 / .3).
Algorithmic bytes per image: read + write of the 3.2 MB image per applied pass (copy 2x, each colour op 2x (+1x for the
contrast mean), grayscale 2x, blur 6 x 2x, erase ~0.1x) — the roofline figure printed is total bytes moved / time."""
import os, random, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unbiased-teacher-v2_b200")]
from ubteacher.data.gpu_augmentation import GpuStrongAugmentation

B, H, W = int(os.environ.get("B", 16)), 800, 1333
g = torch.Generator().manual_seed(0)
imgs = [torch.randint(0, 256, (3, H, W), generator=g, dtype=torch.uint8).cuda() for _ in range(B)]
aug = GpuStrongAugmentation()
torch.manual_seed(0); random.seed(0)
for _ in range(3):
    aug(imgs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters, passes = 20, 0.0
t0 = time.perf_counter(); e0.record()
for _ in range(iters):
    aug(imgs)
    for p in aug.last_params:
        passes += 2 + (2 * 4 + 1 if p["jitter"] else 0) + (2 if p["gray"] else 0) + (12 if p["blur"] else 0)
e1.record(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
dev_s = e0.elapsed_time(e1) / 1e3
img_bytes = 3 * H * W
print(f"device: {B * iters / dev_s:9.1f} images/s (CUDA events; host wall {B * iters / wall:9.1f} images/s incl. the draws), "
      f"{passes * img_bytes / dev_s / 1e9:7.1f} GB/s moved, {dev_s / iters * 1e3:.3f} ms per batch of {B}")
# kernel-only: replay the same parameters (no host draws in the loop)
params = aug.last_params
e0.record()
for _ in range(iters):
    aug(imgs, params=params)
e1.record(); torch.cuda.synchronize()
print(f"device, fixed parameters: {B * iters / (e0.elapsed_time(e1) / 1e3):9.1f} images/s")
# two-crop mapper: decoded 480x640 uint8 HWC images on the host -> (strong, weak) views on the device
try:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util_cfg import fcos_cfg
    from ubteacher.data.dataset_mapper import DatasetMapperTwoCropSeparate
    cfg = fcos_cfg(**{"MODEL.DEVICE": "cuda"})
    mapper = DatasetMapperTwoCropSeparate(cfg, True)
    gh = np.random.default_rng(0)
    dicts = [{"image": torch.from_numpy(gh.integers(0, 256, (480, 640, 3), dtype=np.uint8)).pin_memory(),
              "annotations": [{"bbox": [10.0, 20.0, 300.0, 400.0], "category_id": 1, "iscrowd": 0}]} for _ in range(B)]
    for _ in range(2):
        mapper(dicts)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        mapper(dicts)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"two-crop mapper (480x640 -> shortest edge in [400, 1200], H2D + resize/flip + strong aug): {B * 10 / dt:9.1f} images/s (host wall)")
except Exception as e:  # noqa: BLE001
    print("mapper bench failed:", repr(e))
try:
    from PIL import Image
    import torchvision.transforms as T
    from PIL import ImageFilter

    class GaussianBlur:                       # ubteacher/data/transforms/augmentation_impl.py:7-23
        def __init__(self, sigma): self.sigma = sigma
        def __call__(self, x): return x.filter(ImageFilter.GaussianBlur(radius=random.uniform(self.sigma[0], self.sigma[1])))
    ref = T.Compose([T.RandomApply([T.ColorJitter(0.4, 0.4, 0.4, 0.1)], p=0.8), T.RandomGrayscale(p=0.2),
                     T.RandomApply([GaussianBlur([0.1, 2.0])], p=0.5),
                     T.Compose([T.ToTensor(), T.RandomErasing(p=0.7, scale=(0.05, 0.2), ratio=(0.3, 3.3), value="random"),
                                T.RandomErasing(p=0.5, scale=(0.02, 0.2), ratio=(0.1, 6), value="random"),
                                T.RandomErasing(p=0.3, scale=(0.02, 0.2), ratio=(0.05, 8), value="random"), T.ToPILImage()])])
    pil = [Image.fromarray(im.cpu().numpy().transpose(1, 2, 0), "RGB") for im in imgs[:8]]
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < 10.0:
        for im in pil:
            np.array(ref(im)); n += 1
    dt = time.perf_counter() - t0
    print(f"reference pipeline (PIL / torchvision, 1 host thread): {n / dt:7.1f} images/s  ({os.cpu_count()} cores on this box)")
except Exception as e:  # noqa: BLE001
    print("reference pipeline unavailable:", e)
