from .rpn import PseudoLabRPN  # noqa: F401
