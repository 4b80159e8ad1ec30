"""ORACLE (test infrastructure, never shipped or measured as the product).

fp32 PyTorch restatement of the CTSD-2.1 UNet family: reference
src/dwm/models/crossview_temporal.py (ResBlock :75-164, TemporalBasicTransformerBlock
:167-266, TransformerModel :269-514) and src/dwm/models/crossview_temporal_unet.py
(block classes :10-352, UNetCrossviewTemporalConditionModel :355-835) on top of the
diffusers==0.31.0 pieces they inherit (ResnetBlock2D, TemporalResnetBlock,
BasicTransformerBlock, Downsample2D, Upsample2D, UNetSpatioTemporalConditionModel
members).  Parameter names follow the reference state_dict
(Appendix B).

PINNED AGAINST THE REFERENCE'S OWN CODE for what this file restates of OpenDWM
(ResBlock, TemporalBasicTransformerBlock, TransformerModel, the five block classes and the
model forward): tests/golden/make_reference_golden.py runs the reference's
UNetCrossviewTemporalConditionModel from /root/reference/src on the diffusers name shim
(tests/golden/diffusers_stub) in four configurations and this module reproduces the outputs
bit-exactly (tests/test_reference_golden.py).  The restated diffusers blocks in here
(ResnetBlock2D, TemporalResnetBlock, BasicTransformerBlock, samplers) also BACK that shim,
so their arithmetic stays PARITY UNPINNED (see oracle/d31.py header).
"""
import einops
import torch
import torch.nn.functional as F
from torch import nn

from . import d31
from .ctsd import AlphaBlender, ImageAdapter


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels=None, temb_channels=512, eps=1e-5):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.norm1 = nn.GroupNorm(32, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(32, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) \
            if in_channels != out_channels else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        if temb is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class TemporalResnetBlock(nn.Module):
    def __init__(self, in_channels, out_channels=None, temb_channels=512, eps=1e-5):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.norm1 = nn.GroupNorm(32, in_channels, eps=eps)
        self.conv1 = nn.Conv3d(in_channels, out_channels, (3, 1, 1), padding=(1, 0, 0))
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(32, out_channels, eps=eps)
        self.conv2 = nn.Conv3d(out_channels, out_channels, (3, 1, 1), padding=(1, 0, 0))
        self.conv_shortcut = nn.Conv3d(in_channels, out_channels, 1) \
            if in_channels != out_channels else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))          # GroupNorm over (C/32, T, H, W)
        if temb is not None:
            t = self.time_emb_proj(F.silu(temb))[:, :, :, None, None]
            h = h + t.permute(0, 2, 1, 3, 4)
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class ResBlock(nn.Module):
    """crossview_temporal.py:75-164."""

    def __init__(self, in_channels, out_channels=None, temb_channels=512, eps=1e-5,
                 enable_temporal=True, merge_factor=0.5,
                 merge_strategy="learned_with_images"):
        super().__init__()
        oc = out_channels if out_channels is not None else in_channels
        self.spatial_res_block = ResnetBlock2D(in_channels, out_channels, temb_channels, eps)
        if enable_temporal:
            self.temporal_res_block = TemporalResnetBlock(oc, oc, temb_channels, eps)
            self.time_mixer = AlphaBlender(merge_factor, merge_strategy)
        else:
            self.temporal_res_block = None

    def forward(self, hidden_states, temb=None, disable_temporal=None):
        batch_size = hidden_states.shape[0]
        hidden_states = self.spatial_res_block(
            hidden_states.flatten(0, 2),
            temb.flatten(0, 2) if temb is not None else temb)\
            .unflatten(0, tuple(hidden_states.shape[:3]))
        if self.temporal_res_block is not None:
            th = self.temporal_res_block(
                hidden_states.permute(0, 2, 3, 1, 4, 5).flatten(0, 1),
                temb.transpose(1, 2).flatten(0, 1) if temb is not None else temb)\
                .unflatten(0, (batch_size, -1)).permute(0, 3, 1, 2, 4, 5)
            hidden_states = self.time_mixer(
                hidden_states, th, image_only_indicator=disable_temporal)
        return hidden_states


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, head_dim, cross_attention_dim=None):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = d31.Attention(dim, heads, head_dim, bias=False)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = _CrossAttention(dim, cross_attention_dim, heads, head_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = d31.FeedForward(dim, activation_fn="geglu")

    def forward(self, x, encoder_hidden_states=None):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), encoder_hidden_states) + x
        return self.ff(self.norm3(x)) + x


class _CrossAttention(nn.Module):
    """diffusers Attention with cross_attention_dim (falls back to self-attention when no
    encoder_hidden_states are given, like diffusers)."""

    def __init__(self, query_dim, cross_dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        cross_dim = query_dim if cross_dim is None else cross_dim
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        b = x.shape[0]

        def hd(t):
            return t.view(b, -1, self.heads, self.dim_head).transpose(1, 2)
        o = F.scaled_dot_product_attention(hd(self.to_q(x)), hd(self.to_k(ctx)),
                                           hd(self.to_v(ctx)))
        return self.to_out[0](o.transpose(1, 2).reshape(b, -1, self.heads * self.dim_head))


class TemporalBasicTransformerBlock(nn.Module):
    """crossview_temporal.py:167-266 (cross_attention_dim=None, as TransformerModel uses)."""

    def __init__(self, dim, time_mix_inner_dim, heads, head_dim):
        super().__init__()
        self.is_res = dim == time_mix_inner_dim
        self.norm_in = nn.LayerNorm(dim)
        self.ff_in = d31.FeedForward(dim, dim_out=time_mix_inner_dim, activation_fn="geglu")
        self.norm1 = nn.LayerNorm(time_mix_inner_dim)
        self.attn1 = d31.Attention(time_mix_inner_dim, heads, head_dim, bias=False)
        self.norm3 = nn.LayerNorm(time_mix_inner_dim)
        self.ff = d31.FeedForward(time_mix_inner_dim, activation_fn="geglu")

    def forward(self, hidden_states, num_frames, self_attention_mask=None):
        batch_frames, seq_length, _ = hidden_states.shape
        batch_size = batch_frames // num_frames
        h = hidden_states.unflatten(0, (batch_size, -1)).transpose(1, 2).flatten(0, 1)
        residual = h
        h = self.ff_in(self.norm_in(h))
        if self.is_res:
            h = h + residual
        if self_attention_mask is not None:
            self_attention_mask = self_attention_mask.repeat_interleave(seq_length, 0)
        h = self.attn1(self.norm1(h), attention_mask=self_attention_mask) + h
        ff = self.ff(self.norm3(h))
        h = ff + h if self.is_res else ff
        return h.unflatten(0, (batch_size, -1)).transpose(1, 2).flatten(0, 1)


class TransformerModel(nn.Module):
    """crossview_temporal.py:269-514."""

    def __init__(self, num_attention_heads=16, attention_head_dim=88, in_channels=320,
                 enable_crossview=True, enable_temporal=True,
                 enable_rowwise_crossview=False, enable_rowwise_temporal=False,
                 num_layers=1, cross_attention_dim=None, merge_factor=0.5,
                 merge_strategy="learned_with_images"):
        super().__init__()
        inner = num_attention_heads * attention_head_dim
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner, num_attention_heads, attention_head_dim,
                                  cross_attention_dim) for _ in range(num_layers)])
        self.time_proj = d31.Timesteps(in_channels, True, 0)
        self.enable_rowwise_crossview = enable_rowwise_crossview
        self.enable_rowwise_temporal = enable_rowwise_temporal
        if enable_crossview:
            self.view_pos_embed = d31.TimestepEmbedding(in_channels, in_channels * 4, in_channels)
            self.crossview_transformer_blocks = nn.ModuleList([
                TemporalBasicTransformerBlock(inner, inner, num_attention_heads,
                                              attention_head_dim) for _ in range(num_layers)])
            self.view_mixer = AlphaBlender(merge_factor, merge_strategy)
        else:
            self.view_pos_embed = None
            self.crossview_transformer_blocks = [None] * num_layers
        if enable_temporal:
            self.time_pos_embed = d31.TimestepEmbedding(in_channels, in_channels * 4, in_channels)
            self.temporal_transformer_blocks = nn.ModuleList([
                TemporalBasicTransformerBlock(inner, inner, num_attention_heads,
                                              attention_head_dim) for _ in range(num_layers)])
            self.time_mixer = AlphaBlender(merge_factor, merge_strategy)
        else:
            self.time_pos_embed = None
            self.temporal_transformer_blocks = [None] * num_layers
        self.proj_out = nn.Linear(inner, in_channels)

    def _crossview(self, block, h, view_emb, mask, batch_size, view_count, width, disable):
        c = h + view_emb
        if self.enable_rowwise_crossview:
            c = einops.rearrange(c, "btv (h w) c -> (btv w) h c", w=width)
            c = block(c, num_frames=view_count * width, self_attention_mask=mask)
            c = einops.rearrange(c, "(btv w) h c -> btv (h w) c", w=width)
        else:
            c = block(c, num_frames=view_count, self_attention_mask=mask)
        return self.view_mixer(h.unflatten(0, (batch_size, -1)), c.unflatten(0, (batch_size, -1)),
                               image_only_indicator=disable).flatten(0, 1)

    def _temporal(self, block, h, seq_emb, batch_size, sequence_length, width, disable):
        t = h + seq_emb
        if self.enable_rowwise_temporal:
            t = einops.rearrange(t, "(b t v) (h w) c -> (b v t w) h c", b=batch_size,
                                 t=sequence_length, w=width)
            t = block(t, num_frames=sequence_length * width)
            t = einops.rearrange(t, "(b v t w) h c -> (b t v) (h w) c", b=batch_size,
                                 t=sequence_length, w=width)
        else:
            t = einops.rearrange(t, "(b t v) hw c -> (b v t) hw c", b=batch_size, t=sequence_length)
            t = block(t, num_frames=sequence_length)
            t = einops.rearrange(t, "(b v t) hw c -> (b t v) hw c", b=batch_size, t=sequence_length)
        return self.time_mixer(h.unflatten(0, (batch_size, -1)), t.unflatten(0, (batch_size, -1)),
                               image_only_indicator=disable).flatten(0, 1)

    def forward(self, hidden_states, encoder_hidden_states=None, disable_crossview=None,
                disable_temporal=None, crossview_attention_mask=None):
        B, T, V, _, height, width = hidden_states.shape
        residual = hidden_states
        ctx = encoder_hidden_states.flatten(0, 2) if encoder_hidden_states is not None else None
        h = self.norm(hidden_states.flatten(0, 2))
        h = self.proj_in(h.flatten(2).transpose(-2, -1))
        if self.view_pos_embed is not None:
            ve = torch.arange(V, device=h.device).view(1, 1, V).repeat(B, T, 1)
            ve = self.view_pos_embed(self.time_proj(ve.flatten()).to(h.dtype)).unsqueeze(1)
        if self.time_pos_embed is not None:
            se = torch.arange(T, device=h.device).view(1, T, 1).repeat(B, 1, V)
            se = self.time_pos_embed(self.time_proj(se.flatten()).to(h.dtype)).unsqueeze(1)
        mask = crossview_attention_mask
        if self.enable_rowwise_crossview and mask is not None:
            mask = mask.repeat_interleave(width, 2).repeat_interleave(width, 1)\
                .repeat_interleave(T, 0)
        for blk, cv, tp in zip(self.transformer_blocks, self.crossview_transformer_blocks,
                               self.temporal_transformer_blocks):
            h = blk(h, encoder_hidden_states=ctx)
            if cv is not None:
                h = self._crossview(cv, h, ve, mask, B, V, width, disable_crossview)
            if tp is not None:
                h = self._temporal(tp, h, se, B, T, width, disable_temporal)
        h = self.proj_out(h).transpose(-2, -1).view(B, T, V, -1, height, width).contiguous()
        return h + residual


class Downsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _UBlock(nn.Module):
    """Down / Up / Mid block of crossview_temporal_unet.py:10-352."""

    def __init__(self, kind, res_specs, attn_channels, temb_channels, eps, heads,
                 cross_attention_dim, tlayers, sampler, kw):
        super().__init__()
        self.kind = kind
        self.resnets = nn.ModuleList([
            ResBlock(i, o, temb_channels=temb_channels, eps=eps,
                     enable_temporal=kw["enable_temporal"], merge_factor=kw["merge_factor"])
            for i, o in res_specs])
        n_attn = 0 if attn_channels is None else (len(res_specs) - (1 if kind == "mid" else 0))
        if n_attn:
            self.attentions = nn.ModuleList([
                TransformerModel(heads, attn_channels // heads, in_channels=attn_channels,
                                 enable_crossview=kw["enable_crossview"],
                                 enable_temporal=kw["enable_temporal"],
                                 enable_rowwise_crossview=kw["enable_rowwise_crossview"],
                                 enable_rowwise_temporal=kw["enable_rowwise_temporal"],
                                 num_layers=tlayers, cross_attention_dim=cross_attention_dim,
                                 merge_factor=kw["merge_factor"]) for _ in range(n_attn)])
        else:
            self.attentions = None
        if sampler == "down":
            self.downsamplers = nn.ModuleList([Downsample2D(res_specs[-1][1])])
        elif sampler == "up":
            self.upsamplers = nn.ModuleList([Upsample2D(res_specs[-1][1])])


class UNetCrossviewTemporalConditionModel(nn.Module):
    """crossview_temporal_unet.py:355-835 (depth net / align projection omitted: no
    shipped CTSD config enables them)."""

    def __init__(self, sample_size=None, in_channels=8, out_channels=4,
                 down_block_types=("CrossAttnDownBlockCrossviewTemporal",) * 3 +
                 ("DownBlockCrossviewTemporal",),
                 up_block_types=("UpBlockCrossviewTemporal",) +
                 ("CrossAttnUpBlockCrossviewTemporal",) * 3,
                 block_out_channels=(320, 640, 1280, 1280), addition_time_embed_dim=256,
                 projection_class_embeddings_input_dim=768, layers_per_block=2,
                 norm_eps=1e-5, cross_attention_dim=1024, transformer_layers_per_block=1,
                 num_attention_heads=(5, 10, 20, 20), merge_factor=0.5,
                 enable_crossview=True, enable_temporal=True,
                 enable_rowwise_crossview=False, enable_rowwise_temporal=False,
                 condition_image_adapter_config=None, depth_net_config=None,
                 depth_frustum_range=None, enforce_align_projection=None):
        super().__init__()
        assert depth_net_config is None and enforce_align_projection is None
        n = len(block_out_channels)
        boc = block_out_channels
        temb = boc[0] * 4
        heads = (num_attention_heads,) * n if isinstance(num_attention_heads, int) \
            else tuple(num_attention_heads)
        lpb = [layers_per_block] * n if isinstance(layers_per_block, int) else list(layers_per_block)
        tl = [transformer_layers_per_block] * n if isinstance(transformer_layers_per_block, int) \
            else list(transformer_layers_per_block)
        kw = dict(enable_crossview=enable_crossview, enable_temporal=enable_temporal,
                  enable_rowwise_crossview=enable_rowwise_crossview,
                  enable_rowwise_temporal=enable_rowwise_temporal, merge_factor=merge_factor)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_proj = d31.Timesteps(boc[0], True, 0)
        self.time_embedding = d31.TimestepEmbedding(boc[0], temb)
        self.add_time_proj = d31.Timesteps(addition_time_embed_dim, True, 0)
        self.add_embedding = None if projection_class_embeddings_input_dim is None else \
            d31.TimestepEmbedding(projection_class_embeddings_input_dim, temb)
        self.down_blocks = nn.ModuleList()
        oc = boc[0]
        for i, t in enumerate(down_block_types):
            ic, oc = oc, boc[i]
            specs = [(ic if j == 0 else oc, oc) for j in range(lpb[i])]
            attn = oc if t.startswith("CrossAttn") else None
            self.down_blocks.append(_UBlock("down", specs, attn, temb, norm_eps, heads[i],
                                            cross_attention_dim, tl[i],
                                            "down" if i != n - 1 else None, kw))
        self.mid_block = _UBlock("mid", [(boc[-1], boc[-1])] * 2, boc[-1], temb, norm_eps,
                                 heads[-1], cross_attention_dim, tl[-1], None, kw)
        self.up_blocks = nn.ModuleList()
        rboc, rheads = list(reversed(boc)), list(reversed(heads))
        rlpb, rtl = list(reversed(lpb)), list(reversed(tl))
        oc = rboc[0]
        for i, t in enumerate(up_block_types):
            prev, oc = oc, rboc[i]
            ic = rboc[min(i + 1, n - 1)]
            layers = rlpb[i] + 1
            specs = [((prev if j == 0 else oc) + (ic if j == layers - 1 else oc), oc)
                     for j in range(layers)]
            attn = oc if t.startswith("CrossAttn") else None
            self.up_blocks.append(_UBlock("up", specs, attn, temb, norm_eps, rheads[i],
                                          cross_attention_dim, rtl[i],
                                          "up" if i != n - 1 else None, kw))
        self.conv_norm_out = nn.GroupNorm(32, boc[0], eps=1e-5)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)
        self.condition_image_adapter = None if condition_image_adapter_config is None else \
            ImageAdapter(**condition_image_adapter_config)

    @staticmethod
    def _run(block, sample, emb, ctx, dcv, dtp, mask, skips=None):
        args = dict(encoder_hidden_states=ctx, disable_crossview=dcv, disable_temporal=dtp,
                    crossview_attention_mask=mask)
        outs = ()
        if block.kind == "mid":
            sample = block.resnets[0](sample, emb, disable_temporal=dtp)
            for attn, res in zip(block.attentions, block.resnets[1:]):
                sample = attn(sample, **args)
                sample = res(sample, emb, disable_temporal=dtp)
            return sample, outs
        for j, res in enumerate(block.resnets):
            if skips is not None:
                sample = torch.cat([sample, skips[-1 - j]], dim=-3)
            sample = res(sample, emb, disable_temporal=dtp)
            if block.attentions is not None:
                sample = block.attentions[j](sample, **args)
            outs = outs + (sample,)
        if hasattr(block, "downsamplers"):
            sample = block.downsamplers[0](sample.flatten(0, 2)).unflatten(0, tuple(sample.shape[:3]))
            outs = outs + (sample,)
        if hasattr(block, "upsamplers"):
            sample = block.upsamplers[0](sample.flatten(0, 2)).unflatten(0, tuple(sample.shape[:3]))
        return sample, outs

    def forward(self, sample, timesteps, frustum_bev_residuals=None,
                encoder_hidden_states=None, condition_image_tensor=None,
                disable_crossview=None, disable_temporal=None, crossview_attention_mask=None,
                camera_intrinsics=None, camera_transforms=None, added_time_ids=None,
                camera_intrinsics_norm=None, camera2referego=None, return_dict=False):
        B, T, V = sample.shape[:3]
        t_emb = self.time_proj(timesteps.flatten()).to(sample.dtype)
        emb = self.time_embedding(t_emb).unflatten(0, timesteps.shape[:3])
        if added_time_ids is not None:
            aug = self.add_time_proj(added_time_ids.flatten()).to(sample.dtype)
            emb = emb + self.add_embedding(aug.view(B * T * V, -1)).view(B, T, V, -1)
        residuals = None
        if self.condition_image_adapter is not None and condition_image_tensor is not None:
            residuals = self.condition_image_adapter(condition_image_tensor)
        sample = self.conv_in(sample.flatten(0, 2)).unflatten(0, (B, T, V))
        if residuals:
            sample = sample + residuals.pop(0)
        skips = (sample,)
        for blk in self.down_blocks:
            sample, outs = self._run(blk, sample, emb, encoder_hidden_states,
                                     disable_crossview, disable_temporal,
                                     crossview_attention_mask)
            if residuals:
                sample = sample + residuals.pop(0)
                outs = outs[:-1] + (sample,)
            skips = skips + outs
        sample, _ = self._run(self.mid_block, sample, emb, encoder_hidden_states,
                              disable_crossview, disable_temporal, crossview_attention_mask)
        for blk in self.up_blocks:
            k = len(blk.resnets)
            res, skips = skips[-k:], skips[:-k]
            sample, _ = self._run(blk, sample, emb, encoder_hidden_states, disable_crossview,
                                  disable_temporal, crossview_attention_mask, skips=res)
        sample = self.conv_out(F.silu(self.conv_norm_out(sample.flatten(0, 2))))
        sample = sample.view(B, T, V, *sample.shape[1:])
        if return_dict:
            return {"noise_pred": sample}
        return (sample,)
