from .fcos_outputs import BoxSet, FCOSOutputs  # noqa: F401
