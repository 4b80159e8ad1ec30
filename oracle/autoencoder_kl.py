"""TEST INFRASTRUCTURE ONLY — fp32 PyTorch restatement of the decode path (and the encode
path: vae.py `Encoder`, `DownEncoderBlock2D`, downsampling.py `Downsample2D`,
`DiagonalGaussianDistribution`) of diffusers==0.31.0 `AutoencoderKL` (models/autoencoders/autoencoder_kl.py + vae.py `Decoder`,
unets/unet_2d_blocks.py `UNetMidBlock2D` / `UpDecoderBlock2D`, resnet.py `ResnetBlock2D`,
upsampling.py `Upsample2D`, attention_processor.py `Attention` with `group_norm` and
`residual_connection`), which the reference calls for every emitted frame of configs 2-4
(src/dwm/pipelines/ctsd.py:1633-1640, :2095-2098).

diffusers is not installed in this image and nothing in the reference pins these numerics:
PARITY UNPINNED (restated from the published 0.31.0 source; key names follow the SD-2.1 /
SD-3.5 `vae/diffusion_pytorch_model.safetensors` layout so real checkpoints load).
Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module.
"""
import torch
import torch.nn.functional as F
from torch import nn


class ResnetBlock2D(nn.Module):
    """resnet.py ResnetBlock2D with temb_channels=None, eps 1e-6, output_scale_factor 1."""

    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    """attention_processor.py Attention(heads = C // attention_head_dim, norm_num_groups,
    residual_connection=True, bias=True) run by AttnProcessor2_0 on a [B,C,H,W] input."""

    def __init__(self, channels, head_dim, groups):
        super().__init__()
        self.heads = channels // head_dim
        self.group_norm = nn.GroupNorm(groups, channels, eps=1e-6)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])

    def forward(self, x):
        B, C, H, W = x.shape
        h = x.view(B, C, H * W).transpose(1, 2)
        h = self.group_norm(h.transpose(1, 2)).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        d = C // self.heads
        q, k, v = (t.view(B, -1, self.heads, d).transpose(1, 2) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.transpose(1, 2).reshape(B, -1, C)
        o = self.to_out[0](o)
        return o.transpose(-1, -2).reshape(B, C, H, W) + x


class UNetMidBlock2D(nn.Module):
    def __init__(self, channels, head_dim, groups, add_attention=True):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(channels, channels, groups),
                                      ResnetBlock2D(channels, channels, groups)])
        self.attentions = nn.ModuleList(
            [Attention(channels, head_dim, groups) if add_attention else None])

    def forward(self, x):
        x = self.resnets[0](x)
        for attn, res in zip(self.attentions, self.resnets[1:]):
            if attn is not None:
                x = attn(x)
            x = res(x)
        return x


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class UpDecoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_upsample):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Decoder(nn.Module):
    """vae.py Decoder (norm_type "group")."""

    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block,
                 groups, mid_block_add_attention=True):
        super().__init__()
        rev = list(reversed(block_out_channels))
        self.conv_in = nn.Conv2d(in_channels, rev[0], 3, padding=1)
        self.mid_block = UNetMidBlock2D(rev[0], rev[0], groups, mid_block_add_attention)
        self.up_blocks = nn.ModuleList()
        out = rev[0]
        for i, ch in enumerate(rev):
            prev, out = out, ch
            self.up_blocks.append(UpDecoderBlock2D(prev, out, layers_per_block + 1, groups,
                                                   i != len(rev) - 1))
        self.conv_norm_out = nn.GroupNorm(groups, rev[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(rev[-1], out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Downsample2D(nn.Module):
    """downsampling.py Downsample2D(use_conv=True, padding=0): pad right / bottom by one, then
    a stride-2 3x3 convolution (key `downsamplers.0.conv`)."""

    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class Encoder(nn.Module):
    """vae.py Encoder (double_z: conv_out emits mean | logvar)."""

    def __init__(self, in_channels, latent_channels, block_out_channels, layers_per_block,
                 groups, mid_block_add_attention=True):
        super().__init__()
        boc = list(block_out_channels)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out = boc[0]
        for i, ch in enumerate(boc):
            prev, out = out, ch
            self.down_blocks.append(DownEncoderBlock2D(prev, out, layers_per_block, groups,
                                                       i != len(boc) - 1))
        self.mid_block = UNetMidBlock2D(boc[-1], boc[-1], groups, mid_block_add_attention)
        self.conv_norm_out = nn.GroupNorm(groups, boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], 2 * latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    """vae.py DiagonalGaussianDistribution: moments = mean | logvar along channels."""

    def __init__(self, parameters):
        self.mean, logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator).to(self.mean)
        return self.mean + self.std * noise


class AutoencoderKL(nn.Module):
    """autoencoder_kl.py AutoencoderKL: `encode` = Encoder, quant_conv if configured,
    DiagonalGaussianDistribution; `decode` = post_quant_conv if configured, then Decoder."""

    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(64,),
                 layers_per_block=1, latent_channels=4, norm_num_groups=32,
                 scaling_factor=0.18215, shift_factor=None, use_quant_conv=True,
                 use_post_quant_conv=True, mid_block_add_attention=True, **unused):
        super().__init__()
        # decoder-side modules are registered first: seeded test weights of the decode path do
        # not depend on whether the encoder exists
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels,
                               layers_per_block, norm_num_groups, mid_block_add_attention)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1) \
            if use_post_quant_conv else None
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels,
                               layers_per_block, norm_num_groups, mid_block_add_attention)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1) \
            if use_quant_conv else None

    def encode(self, x):
        h = self.encoder(x)
        if self.quant_conv is not None:
            h = self.quant_conv(h)
        return type("AutoencoderKLOutput", (), {"latent_dist": DiagonalGaussianDistribution(h)})()

    def decode(self, z, return_dict=True):
        if self.post_quant_conv is not None:
            z = self.post_quant_conv(z)
        dec = self.decoder(z)
        return (dec,) if not return_dict else {"sample": dec}
