from oracle.d31 import FeedForward  # noqa: F401


class BasicTransformerBlock:      # UNet family only; pinned through oracle/unet.py
    def __init__(self, *a, **k):
        raise NotImplementedError("BasicTransformerBlock is outside the DiT shim")
