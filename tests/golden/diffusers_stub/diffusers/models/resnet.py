"""diffusers.models.resnet names used by the reference's ResBlock
(/root/reference/src/dwm/models/crossview_temporal.py:104-113) -> oracle/unet.py restatements."""
from oracle.unet import (Downsample2D as _Down, ResnetBlock2D, TemporalResnetBlock,  # noqa: F401
                         Upsample2D as _Up)


class Downsample2D(_Down):
    def __init__(self, channels, use_conv=True, out_channels=None, padding=1, name="conv"):
        assert use_conv and (out_channels is None or out_channels == channels)
        super().__init__(channels)


class Upsample2D(_Up):
    def __init__(self, channels, use_conv=True, out_channels=None):
        assert use_conv and (out_channels is None or out_channels == channels)
        super().__init__(channels)
