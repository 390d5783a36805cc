"""Two-crop semi-supervised train loader (reference: ubteacher/data/build.py:37-61 ``divide_label_unlabel``, :144-272
``build_detection_semisup_train_loader_two_crops`` / ``build_semisup_batch_data_loader_two_crop``), rebuilt around the device
mapper: the host only samples indices, groups the dataset dicts by orientation and decodes files; resize / flip / strong
augmentation run batched on the GPU (data/dataset_mapper.py).

    dataset dicts --TrainingSampler (one shuffled, rank-sharded stream per side)--> AspectRatioGroupedSemiSupDatasetTwoCrop
    (on the dicts' width / height: the weak transform keeps the orientation, so these are the buckets the reference forms
    after mapping) --> decode the batch's files --> DatasetMapperTwoCropSeparate(batch) --> (label_strong, label_weak,
    unlabel_strong, unlabel_weak)

Dataset registration ([D2] DatasetCatalog / ``get_detection_dataset_dicts``) is not rebuilt: pass the dicts."""
import itertools
import json

import numpy as np
import torch

from ..d2compat import comm
from .common import AspectRatioGroupedSemiSupDatasetTwoCrop


def divide_label_unlabel(dataset_dicts, SupPercent, random_data_seed, random_data_seed_path):
    """Labeled / unlabeled split of a dataset by the index lists shipped in ``dataseed/COCO_supervision.txt``
    (json: {"<percent>": {"<seed>": [indices]}}); order inside both parts follows the dataset."""
    with open(random_data_seed_path, "r") as f:
        table = json.load(f)
    chosen = np.array(table[str(SupPercent)][str(random_data_seed)])
    if chosen.shape[0] != int(SupPercent / 100.0 * len(dataset_dicts)):
        raise AssertionError("Number of READ_DATA is mismatched.")
    chosen = set(chosen.tolist())
    label = [d for i, d in enumerate(dataset_dicts) if i in chosen]
    unlabel = [d for i, d in enumerate(dataset_dicts) if i not in chosen]
    return label, unlabel


class TrainingSampler:
    """[D2] TrainingSampler: an infinite stream of indices — concatenated permutations drawn from ONE seed shared by all
    ranks — of which rank r consumes elements r, r + world, r + 2 world, ..."""

    def __init__(self, size, shuffle=True, seed=0, rank=None, world_size=None):
        assert size > 0
        self.size, self.shuffle, self.seed = size, shuffle, int(seed)
        self.rank = comm.get_rank() if rank is None else rank
        self.world_size = comm.get_world_size() if world_size is None else world_size

    def _stream(self):
        g = torch.Generator()
        g.manual_seed(self.seed)
        while True:
            if self.shuffle:
                yield from torch.randperm(self.size, generator=g).tolist()
            else:
                yield from range(self.size)

    def __iter__(self):
        yield from itertools.islice(self._stream(), self.rank, None, self.world_size)


def read_image(file_name, fmt="BGR"):
    """[D2] detection_utils.read_image for the two formats the recipes use: uint8 [h, w, 3] in RGB or BGR order."""
    from PIL import Image
    with Image.open(file_name) as im:
        arr = np.asarray(im.convert("RGB"))
    return np.ascontiguousarray(arr[:, :, ::-1] if fmt == "BGR" else arr)


class TwoCropBatchLoader:
    """Iterable over (label_strong, label_weak, unlabel_strong, unlabel_weak) batches for one rank."""

    def __init__(self, label_dicts, unlabel_dicts, mapper, batch_size_label, batch_size_unlabel, *, aspect_ratio_grouping=True,
                 seed=0, reader=None, image_format="BGR", rank=None, world_size=None):
        self.label_dicts, self.unlabel_dicts, self.mapper = label_dicts, unlabel_dicts, mapper
        self.bl, self.bu = batch_size_label, batch_size_unlabel
        self.grouping = aspect_ratio_grouping
        self.label_sampler = TrainingSampler(len(label_dicts), seed=seed, rank=rank, world_size=world_size)
        self.unlabel_sampler = TrainingSampler(len(unlabel_dicts), seed=seed + 1, rank=rank, world_size=world_size)
        self.reader = reader if reader is not None else (lambda d: read_image(d["file_name"], image_format))

    def _pairs(self, dicts, sampler):
        for i in sampler:
            d = dicts[i]
            yield d, d                 # (strong, weak) slots of the grouping stage: the same dict before mapping

    def _decoded(self, batch):
        out = []
        for d in batch:
            d = dict(d)
            if "image" not in d:
                d["image"] = self.reader(d)
            out.append(d)
        return out

    def __iter__(self):
        lab, unl = self._pairs(self.label_dicts, self.label_sampler), self._pairs(self.unlabel_dicts, self.unlabel_sampler)
        if self.grouping:
            batches = AspectRatioGroupedSemiSupDatasetTwoCrop((lab, unl), (self.bl, self.bu))
        else:
            batches = (([p[0] for p in itertools.islice(lab, self.bl)], None, [p[0] for p in itertools.islice(unl, self.bu)], None)
                       for _ in itertools.count())
        for l_dicts, _, u_dicts, _ in batches:
            lq, lk = self.mapper(self._decoded(l_dicts))
            uq, uk = self.mapper(self._decoded(u_dicts))
            yield lq, lk, uq, uk


def build_detection_semisup_train_loader_two_crops(cfg, mapper=None, dataset_dicts=None, label_dicts=None, unlabel_dicts=None,
                                                   reader=None):
    """build.py:144-272. ``dataset_dicts`` (split by DATALOADER.SUP_PERCENT / RANDOM_DATA_SEED) or the two parts directly
    (DATASETS.CROSS_DATASET). Global batch sizes are divided by the world size like the reference does."""
    if label_dicts is None or unlabel_dicts is None:
        if dataset_dicts is None:
            raise ValueError("pass dataset_dicts (or label_dicts and unlabel_dicts): dataset registration is not rebuilt")
        label_dicts, unlabel_dicts = divide_label_unlabel(dataset_dicts, cfg.DATALOADER.SUP_PERCENT, cfg.DATALOADER.RANDOM_DATA_SEED,
                                                          cfg.DATALOADER.RANDOM_DATA_SEED_PATH)
    if mapper is None:
        from .dataset_mapper import DatasetMapperTwoCropSeparate
        mapper = DatasetMapperTwoCropSeparate(cfg, True)
    name = cfg.DATALOADER.SAMPLER_TRAIN
    if name == "RepeatFactorTrainingSampler":
        raise NotImplementedError("{} not yet supported.".format(name))
    if name != "TrainingSampler":
        raise ValueError("Unknown training sampler: {}".format(name))
    world = comm.get_world_size()
    tl, tu = cfg.SOLVER.IMG_PER_BATCH_LABEL, cfg.SOLVER.IMG_PER_BATCH_UNLABEL
    assert tl > 0 and tl % world == 0, f"Total label batch size ({tl}) must be divisible by the number of gpus ({world})."
    assert tu > 0 and tu % world == 0, f"Total unlabel batch size ({tu}) must be divisible by the number of gpus ({world})."
    return TwoCropBatchLoader(label_dicts, unlabel_dicts, mapper, tl // world, tu // world,
                              aspect_ratio_grouping=cfg.DATALOADER.ASPECT_RATIO_GROUPING, seed=max(int(cfg.SEED), 0), reader=reader,
                              image_format=cfg.INPUT.FORMAT)
