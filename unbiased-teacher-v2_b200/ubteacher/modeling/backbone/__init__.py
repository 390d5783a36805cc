from .fpn import LastLevelP6P7, build_fcos_resnet_fpn_backbone  # noqa: F401
