"""Synthetic two-crop batches with the shape of the reference loader's output
(ubteacher/data/build.py:144-272 yields (label_strong, label_weak, unlabel_strong, unlabel_weak) lists of
dicts): uint8 BGR images 3x800x1333 in pinned host memory, ~7 COCO-like boxes per labeled image
(SURVEY.md §8d). Used by bench.py, smoke() and the tests — there is no dataset on the box."""
import torch

from ..d2compat.structures import Boxes, Instances


def synth_instances(g, h, w, n=7):
    wh = torch.exp(torch.rand(n, 2, generator=g) * (6.238 - 3.466) + 3.466)       # log-uniform [32, 512]
    wh[:, 0].clamp_(max=w - 2.0)
    wh[:, 1].clamp_(max=h - 2.0)
    xy = torch.rand(n, 2, generator=g) * (torch.tensor([float(w), float(h)]) - wh)
    inst = Instances((h, w))
    inst.gt_boxes = Boxes(torch.cat([xy, xy + wh], 1))
    inst.gt_classes = torch.randint(0, 80, (n,), generator=g)
    return inst


class SyntheticTwoCropLoader:
    """device=None: images live in pinned host memory and GT are host ``Instances`` (the end-to-end mode, the
    model does the H2D copies); device="cuda": the image pool and pre-packed GT ``BoxSet``s are resident in HBM."""

    def __init__(self, n_label, n_unlabel, h=800, w=1333, rank=0, boxes_per_image=7, pool=2, pin=True, device=None,
                 strong_augment=False):
        self.nl, self.nu, self.h, self.w, self.rank, self.nbox = n_label, n_unlabel, h, w, rank, boxes_per_image
        self.device = device
        # strong_augment: the strong ("q") views are produced from the weak ("k") views by the device two-crop
        # augmentation (dataset_mapper.py:116-131: image_strong_aug = strong_augmentation(image_weak_aug)), every step
        self.aug = None
        if strong_augment:
            from .gpu_augmentation import GpuStrongAugmentation
            self.aug = GpuStrongAugmentation()
        g = torch.Generator().manual_seed(20260 + 1000 * rank)
        # a small pool of random images, re-used round-robin (generating 32 fresh 3.2 MB images per step on the
        # host would measure torch.randint, not the training step)
        n_img = pool * 2 * (n_label + n_unlabel)
        self.pool = []
        for _ in range(n_img):
            t = torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8)
            if device is not None:
                t = t.to(device)
            elif pin and torch.cuda.is_available():
                t = t.pin_memory()
            self.pool.append(t)
        self.g = g
        self.step = 0
        self.gt_pool = []
        if device is not None:
            from ..modeling.fcos.fcos_outputs import BoxSet
            for _ in range(4):
                insts = [synth_instances(g, h, w, boxes_per_image) for _ in range(n_label)]
                self.gt_pool.append(BoxSet.from_instances(insts + insts, device))
            torch.cuda.synchronize()

    def __iter__(self):
        return self

    def _img(self):
        t = self.pool[self.step % len(self.pool)]
        self.step += 1
        return t

    def __next__(self):
        if self.device is not None:
            mk = lambda n: [{"image": self._img(), "height": self.h, "width": self.w} for _ in range(n)]
            lq, lk, uq, uk = mk(self.nl), mk(self.nl), mk(self.nu), mk(self.nu)
            gt = self.gt_pool[(self.step // max(len(self.pool), 1)) % len(self.gt_pool)]
            for d in lq + lk:
                d["instances"] = gt          # batch-level device-resident ground truth (strong + weak share boxes)
            return self._strong(lq, lk, uq, uk)

        def lab(n):
            q, k = [], []
            for _ in range(n):
                inst = synth_instances(self.g, self.h, self.w, self.nbox)
                q.append({"image": self._img(), "instances": inst, "height": self.h, "width": self.w})
                k.append({"image": self._img(), "instances": inst, "height": self.h, "width": self.w})
            return q, k
        lq, lk = lab(self.nl)
        uq, uk = lab(self.nu)
        return self._strong(lq, lk, uq, uk)

    def _strong(self, lq, lk, uq, uk):
        if self.aug is not None:
            dev = self.device if self.device is not None else torch.device("cuda")
            weak = [d["image"].to(dev, non_blocking=True) for d in lk + uk]       # H2D of the weak views (end-to-end mode)
            for d, w_ in zip(lk + uk, weak):
                d["image"] = w_
            for d, s_ in zip(lq + uq, self.aug(weak)):
                d["image"] = s_
        return lq, lk, uq, uk
