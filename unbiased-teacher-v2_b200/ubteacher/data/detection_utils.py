"""``build_strong_augmentation`` with the reference's name and role (ubteacher/data/detection_utils.py:8-46): returns the
strong-augmentation callable of the two-crop pipeline — here the batched device implementation."""
from .gpu_augmentation import GpuStrongAugmentation


def build_strong_augmentation(cfg, is_train):
    return GpuStrongAugmentation(is_train=is_train)
