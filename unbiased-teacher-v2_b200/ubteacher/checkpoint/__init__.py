from .detection_checkpoint import DetectionTSCheckpointer, convert_c2_resnet_names  # noqa: F401
