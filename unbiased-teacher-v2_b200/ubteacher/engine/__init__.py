"""Trainers of the B200 hot path under the reference's names (``ubteacher.engine.UBTeacherTrainer`` for FCOS,
``UBRCNNTeacherTrainer`` for Faster R-CNN); see engine/trainer.py."""
from . import trainer as _trainer

UBTeacherTrainer = _trainer.UBTeacherTrainer
UBRCNNTeacherTrainer = _trainer.UBRCNNTeacherTrainer

__all__ = ["UBTeacherTrainer", "UBRCNNTeacherTrainer"]
