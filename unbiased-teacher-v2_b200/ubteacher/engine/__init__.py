from .trainer import UBRCNNTeacherTrainer, UBTeacherTrainer  # noqa: F401
