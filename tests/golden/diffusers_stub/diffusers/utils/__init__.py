from . import torch_utils  # noqa: F401
