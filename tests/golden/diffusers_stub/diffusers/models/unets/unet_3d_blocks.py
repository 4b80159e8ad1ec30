"""Base classes of the reference's UNet blocks
(/root/reference/src/dwm/models/crossview_temporal_unet.py:10-352).  The reference calls
their constructors positionally and then REPLACES `resnets` / `attentions` with its own
modules, so only what survives that is built here: the down / up samplers, the
`has_cross_attention` flag and `gradient_checkpointing`."""
from torch import nn

from ..resnet import Downsample2D, Upsample2D


class _Base(nn.Module):
    def __init__(self):
        super().__init__()
        self.gradient_checkpointing = False


class UNetMidBlockSpatioTemporal(_Base):
    def __init__(self, in_channels, temb_channels, num_layers=1, transformer_layers_per_block=1,
                 num_attention_heads=1, cross_attention_dim=1280):
        super().__init__()
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads


class DownBlockSpatioTemporal(_Base):
    def __init__(self, in_channels, out_channels, temb_channels, num_layers=1,
                 add_downsample=True):
        super().__init__()
        self.downsamplers = nn.ModuleList(
            [Downsample2D(out_channels, use_conv=True, out_channels=out_channels, name="op")]) \
            if add_downsample else None


class CrossAttnDownBlockSpatioTemporal(_Base):
    def __init__(self, in_channels, out_channels, temb_channels, num_layers=1,
                 transformer_layers_per_block=1, num_attention_heads=1,
                 cross_attention_dim=1280, add_downsample=True):
        super().__init__()
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        self.downsamplers = nn.ModuleList(
            [Downsample2D(out_channels, use_conv=True, out_channels=out_channels, name="op")]) \
            if add_downsample else None


class UpBlockSpatioTemporal(_Base):
    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels,
                 resolution_idx=None, num_layers=1, resnet_eps=1e-6, add_upsample=True):
        super().__init__()
        self.resolution_idx = resolution_idx
        self.upsamplers = nn.ModuleList(
            [Upsample2D(out_channels, use_conv=True, out_channels=out_channels)]) \
            if add_upsample else None


class CrossAttnUpBlockSpatioTemporal(_Base):
    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels,
                 resolution_idx=None, num_layers=1, transformer_layers_per_block=1,
                 resnet_eps=1e-6, num_attention_heads=1, cross_attention_dim=1280,
                 add_upsample=True):
        super().__init__()
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        self.resolution_idx = resolution_idx
        self.upsamplers = nn.ModuleList(
            [Upsample2D(out_channels, use_conv=True, out_channels=out_channels)]) \
            if add_upsample else None


def get_down_block(*a, **k):
    raise NotImplementedError("stock SVD down blocks are not used by any CTSD config")


def get_up_block(*a, **k):
    raise NotImplementedError("stock SVD up blocks are not used by any CTSD config")
