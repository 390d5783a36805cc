from .trainer import UBTeacherTrainer  # noqa: F401
