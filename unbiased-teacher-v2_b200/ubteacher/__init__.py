"""B200-native Unbiased Teacher v2 hot path behind the reference's Python surface
(``ubteacher.add_ubteacher_config``, ``ubteacher.engine.UBTeacherTrainer`` ...)."""
from .config import add_ubteacher_config  # noqa: F401
