"""ORACLE (test infrastructure, never shipped or measured as the product).

fp32 PyTorch restatement of the DECODER of diffusers==0.31.0 `AutoencoderKLCogVideoX`
(THUDM/CogVideoX-2b `vae/`), the temporal VAE that config 5 of the reference selects
(`examples/ctsd_35_tvae_6views_video_generation_with_layout.json:51-52`) and that
`CrossviewTemporalSD.inference_pipeline` calls at src/dwm/pipelines/ctsd.py:1628-1643
(DF variant :1609-1621).  Follows SURVEY.md Appendix A.7; parameter names follow the
diffusers state_dict (`decoder.*`).  PARITY UNPINNED (see oracle/d31.py header).
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn


class CausalConv3d(nn.Module):
    """Temporal left pad by replicating the first frame k-1 times, or by the cached last
    k-1 input frames of the previous chunk; spatial zero pad k//2."""

    def __init__(self, in_channels, out_channels, kernel_size):
        super().__init__()
        self.time_kernel_size = kernel_size
        self.pad = kernel_size // 2
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size)

    def forward(self, inputs, conv_cache=None):
        k = self.time_kernel_size
        if k > 1:
            cached = [conv_cache] if conv_cache is not None else \
                [inputs[:, :, :1]] * (k - 1)
            inputs = torch.cat(cached + [inputs], dim=2)
        new_cache = inputs[:, :, -k + 1:].clone() if k > 1 else None
        p = self.pad
        inputs = F.pad(inputs, (p, p, p, p), mode="constant", value=0)
        return self.conv(inputs), new_cache


class SpatialNorm3D(nn.Module):
    def __init__(self, f_channels, zq_channels, groups=32):
        super().__init__()
        self.norm_layer = nn.GroupNorm(groups, f_channels, eps=1e-6, affine=True)
        self.conv_y = CausalConv3d(zq_channels, f_channels, 1)
        self.conv_b = CausalConv3d(zq_channels, f_channels, 1)

    def forward(self, f, zq):
        if f.shape[2] > 1 and f.shape[2] % 2 == 1:
            z_first = F.interpolate(zq[:, :, :1], size=f[:, :, :1].shape[-3:])
            z_rest = F.interpolate(zq[:, :, 1:], size=f[:, :, 1:].shape[-3:])
            zq = torch.cat([z_first, z_rest], dim=2)
        else:
            zq = F.interpolate(zq, size=f.shape[-3:])
        return self.norm_layer(f) * self.conv_y(zq)[0] + self.conv_b(zq)[0]


class ResnetBlock3D(nn.Module):
    def __init__(self, in_channels, out_channels, spatial_norm_dim, groups=32):
        super().__init__()
        self.norm1 = SpatialNorm3D(in_channels, spatial_norm_dim, groups)
        self.conv1 = CausalConv3d(in_channels, out_channels, 3)
        self.norm2 = SpatialNorm3D(out_channels, spatial_norm_dim, groups)
        self.conv2 = CausalConv3d(out_channels, out_channels, 3)
        self.conv_shortcut = nn.Conv3d(in_channels, out_channels, 1) \
            if in_channels != out_channels else None

    def forward(self, inputs, zq, cache):
        new_cache = {}
        h = F.silu(self.norm1(inputs, zq))
        h, new_cache["conv1"] = self.conv1(h, cache.get("conv1"))
        h = F.silu(self.norm2(h, zq))
        h, new_cache["conv2"] = self.conv2(h, cache.get("conv2"))
        if self.conv_shortcut is not None:
            inputs = self.conv_shortcut(inputs)
        return h + inputs, new_cache


class Upsample3D(nn.Module):
    def __init__(self, channels, compress_time):
        super().__init__()
        self.compress_time = compress_time
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        if self.compress_time:
            if x.shape[2] > 1 and x.shape[2] % 2 == 1:
                x_first = F.interpolate(x[:, :, 0], scale_factor=2.0)
                x_rest = F.interpolate(x[:, :, 1:], scale_factor=2.0)
                x = torch.cat([x_first[:, :, None], x_rest], dim=2)
            elif x.shape[2] > 1:
                x = F.interpolate(x, scale_factor=2.0)
            else:
                x = F.interpolate(x.squeeze(2), scale_factor=2.0)[:, :, None]
        else:
            b, c, t, h, w = x.shape
            x = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
            x = F.interpolate(x, scale_factor=2.0)
            x = x.reshape(b, t, c, *x.shape[2:]).permute(0, 2, 1, 3, 4)
        b, c, t, h, w = x.shape
        x = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
        x = self.conv(x)
        return x.reshape(b, t, *x.shape[1:]).permute(0, 2, 1, 3, 4)


class _Block(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, zq_dim, upsample,
                 compress_time, groups):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock3D(in_channels if i == 0 else out_channels, out_channels,
                          zq_dim, groups) for i in range(num_layers)])
        self.upsamplers = nn.ModuleList(
            [Upsample3D(out_channels, compress_time)]) if upsample else None

    def forward(self, h, zq, cache):
        new_cache = {}
        for i, r in enumerate(self.resnets):
            h, new_cache[i] = r(h, zq, cache.get(i, {}))
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h, new_cache


class Decoder3D(nn.Module):
    def __init__(self, in_channels=16, out_channels=3,
                 block_out_channels=(128, 256, 256, 512), layers_per_block=3,
                 norm_num_groups=32, temporal_compression_ratio=4):
        super().__init__()
        rev = list(reversed(block_out_channels))
        self.conv_in = CausalConv3d(in_channels, rev[0], 3)
        self.mid_block = _Block(rev[0], rev[0], 2, in_channels, False, False,
                                norm_num_groups)
        level = int(np.log2(temporal_compression_ratio))
        self.up_blocks = nn.ModuleList()
        out_ch = rev[0]
        for i in range(len(rev)):
            prev, out_ch = out_ch, rev[i]
            self.up_blocks.append(_Block(
                prev, out_ch, layers_per_block + 1, in_channels,
                i != len(rev) - 1, i < level, norm_num_groups))
        self.norm_out = SpatialNorm3D(rev[-1], in_channels, norm_num_groups)
        self.conv_out = CausalConv3d(rev[-1], out_channels, 3)

    def forward(self, sample, cache=None):
        cache = cache or {}
        new_cache = {}
        h, new_cache["conv_in"] = self.conv_in(sample, cache.get("conv_in"))
        h, new_cache["mid"] = self.mid_block(h, sample, cache.get("mid", {}))
        for i, blk in enumerate(self.up_blocks):
            h, new_cache[i] = blk(h, sample, cache.get(i, {}))
        h = F.silu(self.norm_out(h, sample))
        h, new_cache["conv_out"] = self.conv_out(h, cache.get("conv_out"))
        return h, new_cache


class AutoencoderKLCogVideoXDecoder(nn.Module):
    """decode(z): chunks of `num_latent_frames_batch_size` latent frames (the first chunk
    takes the remainder), causal-conv caches carried across chunks (so GroupNorm
    statistics are per chunk)."""

    scaling_factor = 1.15258426

    def __init__(self, **decoder_kwargs):
        super().__init__()
        self.decoder = Decoder3D(**decoder_kwargs)
        self.num_latent_frames_batch_size = 2

    def decode(self, z):
        fb = self.num_latent_frames_batch_size
        num_frames = z.shape[2]
        cache = None
        out = []
        for i in range(max(num_frames // fb, 1)):
            rem = num_frames % fb
            start = fb * i + (0 if i == 0 else rem)
            end = fb * (i + 1) + rem
            y, cache = self.decoder(z[:, :, start:end], cache)
            out.append(y)
        return torch.cat(out, dim=2)


# -- encoder (diffusers==0.31.0 CogVideoXEncoder3D, AutoencoderKLCogVideoX._encode; reference use:
#    src/dwm/pipelines/ctsd.py:1677-1700 when generate_frames_for_reference is false) ---------------

class EncResnetBlock3D(nn.Module):
    """CogVideoXResnetBlock3D with spatial_norm_dim=None: plain GroupNorm (eps 1e-6)."""

    def __init__(self, in_channels, out_channels, groups=32):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=1e-6)
        self.conv1 = CausalConv3d(in_channels, out_channels, 3)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=1e-6)
        self.conv2 = CausalConv3d(out_channels, out_channels, 3)
        self.conv_shortcut = nn.Conv3d(in_channels, out_channels, 1) \
            if in_channels != out_channels else None

    def forward(self, inputs, cache):
        new_cache = {}
        h, new_cache["conv1"] = self.conv1(F.silu(self.norm1(inputs)), cache.get("conv1"))
        h, new_cache["conv2"] = self.conv2(F.silu(self.norm2(h)), cache.get("conv2"))
        if self.conv_shortcut is not None:
            inputs = self.conv_shortcut(inputs)
        return h + inputs, new_cache


class Downsample3D(nn.Module):
    """CogVideoXDownsample3D: optional temporal average pooling in pairs (an odd frame count
    keeps its first frame), then right / bottom zero pad + stride-2 3x3 conv per frame."""

    def __init__(self, channels, compress_time):
        super().__init__()
        self.compress_time = compress_time
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=0)

    def forward(self, x):
        if self.compress_time:
            b, c, f, h, w = x.shape
            x = x.permute(0, 3, 4, 1, 2).reshape(b * h * w, c, f)
            if f % 2 == 1:
                first, rest = x[..., 0], x[..., 1:]
                if rest.shape[-1] > 0:
                    rest = F.avg_pool1d(rest, kernel_size=2, stride=2)
                x = torch.cat([first[..., None], rest], dim=-1)
            else:
                x = F.avg_pool1d(x, kernel_size=2, stride=2)
            x = x.reshape(b, h, w, c, x.shape[-1]).permute(0, 3, 4, 1, 2)
        x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
        b, c, f, h, w = x.shape
        x = self.conv(x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w))
        return x.reshape(b, f, *x.shape[1:]).permute(0, 2, 1, 3, 4)


class _DownBlock(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, downsample, compress_time, groups):
        super().__init__()
        self.resnets = nn.ModuleList([
            EncResnetBlock3D(in_channels if i == 0 else out_channels, out_channels, groups)
            for i in range(num_layers)])
        self.downsamplers = nn.ModuleList(
            [Downsample3D(out_channels, compress_time)]) if downsample else None

    def forward(self, h, cache):
        new_cache = {}
        for i, r in enumerate(self.resnets):
            h, new_cache[i] = r(h, cache.get(i, {}))
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
        return h, new_cache


class Encoder3D(nn.Module):
    def __init__(self, in_channels=3, out_channels=16,
                 block_out_channels=(128, 256, 256, 512), layers_per_block=3,
                 norm_num_groups=32, temporal_compression_ratio=4):
        super().__init__()
        boc = list(block_out_channels)
        self.conv_in = CausalConv3d(in_channels, boc[0], 3)
        level = int(np.log2(temporal_compression_ratio))
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i in range(len(boc)):
            prev, out_ch = out_ch, boc[i]
            self.down_blocks.append(_DownBlock(prev, out_ch, layers_per_block,
                                               i != len(boc) - 1, i < level, norm_num_groups))
        self.mid_block = _DownBlock(boc[-1], boc[-1], 2, False, False, norm_num_groups)
        self.norm_out = nn.GroupNorm(norm_num_groups, boc[-1], eps=1e-6)
        self.conv_out = CausalConv3d(boc[-1], 2 * out_channels, 3)

    def forward(self, sample, cache=None):
        cache = cache or {}
        new_cache = {}
        h, new_cache["conv_in"] = self.conv_in(sample, cache.get("conv_in"))
        for i, blk in enumerate(self.down_blocks):
            h, new_cache[i] = blk(h, cache.get(i, {}))
        h, new_cache["mid"] = self.mid_block(h, cache.get("mid", {}))
        h, new_cache["conv_out"] = self.conv_out(F.silu(self.norm_out(h)), cache.get("conv_out"))
        return h, new_cache


class AutoencoderKLCogVideoXEncoder(nn.Module):
    """encode(x): chunks of `num_sample_frames_batch_size` (8) frames, the first chunk takes
    the remainder; causal-conv caches carried across chunks; moments = mean | logvar."""

    def __init__(self, **encoder_kwargs):
        super().__init__()
        self.encoder = Encoder3D(**encoder_kwargs)
        self.num_sample_frames_batch_size = 8

    def encode_moments(self, x):
        fb = self.num_sample_frames_batch_size
        n = x.shape[2]
        cache, out = None, []
        for i in range(max(n // fb, 1)):
            rem = n % fb
            start = fb * i + (0 if i == 0 else rem)
            end = fb * (i + 1) + rem
            y, cache = self.encoder(x[:, :, start:end], cache)
            out.append(y)
        return torch.cat(out, dim=2)

    def encode_mode(self, x):
        return torch.chunk(self.encode_moments(x), 2, dim=1)[0]
