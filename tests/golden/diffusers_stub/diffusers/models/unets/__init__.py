from . import unet_3d_blocks  # noqa: F401
