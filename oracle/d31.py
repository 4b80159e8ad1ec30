"""ORACLE (test infrastructure, never shipped or measured as the product).

fp32 PyTorch restatement of the pieces of `diffusers==0.31.0` that OpenDWM's CTSD
path executes.  diffusers is pinned by the reference (requirements.txt:6) but is
neither vendored under /root/reference nor installed here, so these classes
restate the published algorithm of that release and are
anchored on the reference's own call sites (cited per class).

PARITY UNPINNED (tools/pin_diffusers.py pins it the moment diffusers==0.31.0 is importable:
it re-runs the reference on the real package and diffs every golden): the reference ships no
tests, golden vectors or fixtures for
this path and cannot be imported here, so this oracle is
checked for internal consistency only.  Parameter names follow the diffusers
state_dict keys so real checkpoints load unchanged.  (These classes
also back the `diffusers` name shim under tests/golden/diffusers_stub, on which the
reference's OWN modules are imported and run to pin oracle/ctsd.py — that pins the OpenDWM
code paths, not the arithmetic of this file.)
"""
import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn


# -- embeddings ----------------------------------------------------------------

def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False,
                           downscale_freq_shift=1.0, scale=1.0,
                           max_period=10000):
    """diffusers.models.embeddings.get_timestep_embedding (A.1)."""
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(
        0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class Timesteps(nn.Module):
    """Used as index_proj / view_cam_proj (crossview_temporal_dit.py:153-154,163-164)."""

    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift,
                 scale=1):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift
        self.scale = scale

    def forward(self, timesteps):
        return get_timestep_embedding(
            timesteps, self.num_channels, self.flip_sin_to_cos,
            self.downscale_freq_shift, self.scale)


class TimestepEmbedding(nn.Module):
    """linear_2(SiLU(linear_1(x))) (crossview_temporal_dit.py:165-167,174,198)."""

    def __init__(self, in_channels, time_embed_dim, out_dim=None):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(
            time_embed_dim, out_dim if out_dim is not None else time_embed_dim)

    def forward(self, sample):
        return self.linear_2(self.act(self.linear_1(sample)))


class PixArtAlphaTextProjection(nn.Module):
    def __init__(self, in_features, hidden_size, out_features=None):
        super().__init__()
        out_features = hidden_size if out_features is None else out_features
        self.linear_1 = nn.Linear(in_features, hidden_size)
        self.act_1 = nn.SiLU()
        self.linear_2 = nn.Linear(hidden_size, out_features)

    def forward(self, caption):
        return self.linear_2(self.act_1(self.linear_1(caption)))


class CombinedTimestepTextProjEmbeddings(nn.Module):
    """time_text_embed of SD3Transformer2DModel (crossview_temporal_dit.py:431; A.2)."""

    def __init__(self, embedding_dim, pooled_projection_dim):
        super().__init__()
        self.time_proj = Timesteps(256, True, 0)
        self.timestep_embedder = TimestepEmbedding(256, embedding_dim)
        self.text_embedder = PixArtAlphaTextProjection(
            pooled_projection_dim, embedding_dim)

    def forward(self, timestep, pooled_projection):
        t = self.time_proj(timestep)
        t = self.timestep_embedder(t.to(dtype=pooled_projection.dtype))
        return t + self.text_embedder(pooled_projection)


def _sincos_1d(embed_dim, pos):
    omega = np.arange(embed_dim // 2, dtype=np.float64)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def get_2d_sincos_pos_embed(embed_dim, grid_size, base_size=16,
                            interpolation_scale=1.0):
    gh = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size) / \
        interpolation_scale
    gw = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size) / \
        interpolation_scale
    grid = np.stack(np.meshgrid(gw, gh), axis=0)
    grid = grid.reshape([2, 1, grid_size, grid_size])
    emb_h = _sincos_1d(embed_dim // 2, grid[0])
    emb_w = _sincos_1d(embed_dim // 2, grid[1])
    return np.concatenate([emb_h, emb_w], axis=1)


class PatchEmbed(nn.Module):
    """pos_embed of SD3Transformer2DModel (crossview_temporal_dit.py:421; A.3)."""

    def __init__(self, height, width, patch_size, in_channels, embed_dim,
                 pos_embed_max_size):
        super().__init__()
        self.patch_size = patch_size
        self.pos_embed_max_size = pos_embed_max_size
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=patch_size,
                              stride=patch_size, bias=True)
        self.base_size = height // patch_size
        pe = get_2d_sincos_pos_embed(
            embed_dim, pos_embed_max_size, base_size=self.base_size)
        self.register_buffer(
            "pos_embed", torch.from_numpy(pe).float().unsqueeze(0),
            persistent=True)

    def cropped_pos_embed(self, height, width):
        height = height // self.patch_size
        width = width // self.patch_size
        top = (self.pos_embed_max_size - height) // 2
        left = (self.pos_embed_max_size - width) // 2
        pe = self.pos_embed.reshape(
            1, self.pos_embed_max_size, self.pos_embed_max_size, -1)
        pe = pe[:, top:top + height, left:left + width, :]
        return pe.reshape(1, -1, pe.shape[-1])

    def forward(self, latent):
        height, width = latent.shape[-2:]
        latent = self.proj(latent).flatten(2).transpose(1, 2)
        return (latent + self.cropped_pos_embed(height, width)).to(latent.dtype)


# -- norms ---------------------------------------------------------------------

class RMSNorm(nn.Module):
    def __init__(self, dim, eps):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        dt = x.dtype
        var = x.float().pow(2).mean(-1, keepdim=True)
        x = x.float() * torch.rsqrt(var + self.eps)
        return (x * self.weight.float()).to(dt)


class AdaLayerNormZero(nn.Module):
    """chunk order (shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp) (A.4)."""

    def __init__(self, dim):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(dim, 6 * dim)
        self.norm = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)

    def forward(self, x, emb):
        emb = self.linear(self.silu(emb))
        sm, cm, gm, sp, cp, gp = emb.chunk(6, dim=1)
        x = self.norm(x) * (1 + cm[:, None]) + sm[:, None]
        return x, gm, sp, cp, gp


class SD35AdaLayerNormZeroX(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(dim, 9 * dim)
        self.norm = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)

    def forward(self, x, emb):
        emb = self.linear(self.silu(emb))
        sm, cm, gm, sp, cp, gp, sm2, cm2, gm2 = emb.chunk(9, dim=1)
        n = self.norm(x)
        x1 = n * (1 + cm[:, None]) + sm[:, None]
        x2 = n * (1 + cm2[:, None]) + sm2[:, None]
        return x1, gm, sp, cp, gp, x2, gm2


class AdaLayerNormContinuous(nn.Module):
    """chunk order (scale, shift) (A.4)."""

    def __init__(self, dim, cond_dim):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(cond_dim, 2 * dim)
        self.norm = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)

    def forward(self, x, cond):
        emb = self.linear(self.silu(cond).to(x.dtype))
        scale, shift = emb.chunk(2, dim=1)
        return self.norm(x) * (1 + scale)[:, None, :] + shift[:, None, :]


# -- feed forward --------------------------------------------------------------

class GELU(nn.Module):
    def __init__(self, dim_in, dim_out, approximate="none"):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)
        self.approximate = approximate

    def forward(self, x):
        return F.gelu(self.proj(x), approximate=self.approximate)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    """diffusers.models.attention.FeedForward (crossview_temporal.py:548,559; A.6)."""

    def __init__(self, dim, dim_out=None, mult=4, activation_fn="geglu"):
        super().__init__()
        inner = int(dim * mult)
        dim_out = dim if dim_out is None else dim_out
        if activation_fn == "geglu":
            act = GEGLU(dim, inner)
        elif activation_fn == "gelu-approximate":
            act = GELU(dim, inner, approximate="tanh")
        elif activation_fn == "gelu":
            act = GELU(dim, inner)
        else:
            raise ValueError(activation_fn)
        self.net = nn.ModuleList([act, nn.Dropout(0.0), nn.Linear(inner, dim_out)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


# -- attention -----------------------------------------------------------------

class Attention(nn.Module):
    """diffusers.models.attention_processor.Attention with AttnProcessor2_0 /
    JointAttnProcessor2_0 semantics (crossview_temporal.py:552-555; A.5)."""

    def __init__(self, query_dim, heads, dim_head, bias=False, qk_norm=None,
                 eps=1e-5, added_kv_proj_dim=None, context_pre_only=None,
                 out_dim=None):
        super().__init__()
        inner = out_dim if out_dim is not None else dim_head * heads
        self.heads = inner // dim_head if out_dim is not None else heads
        self.dim_head = dim_head
        self.scale = dim_head ** -0.5
        self.context_pre_only = context_pre_only
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(query_dim, inner, bias=bias)
        self.to_v = nn.Linear(query_dim, inner, bias=bias)
        if qk_norm == "rms_norm":
            self.norm_q = RMSNorm(dim_head, eps)
            self.norm_k = RMSNorm(dim_head, eps)
        elif qk_norm is None:
            self.norm_q = self.norm_k = None
        else:
            raise ValueError(qk_norm)
        self.added_kv_proj_dim = added_kv_proj_dim
        if added_kv_proj_dim is not None:
            self.add_k_proj = nn.Linear(added_kv_proj_dim, inner, bias=True)
            self.add_v_proj = nn.Linear(added_kv_proj_dim, inner, bias=True)
            if context_pre_only is not None:
                self.add_q_proj = nn.Linear(added_kv_proj_dim, inner, bias=True)
            if qk_norm == "rms_norm":
                self.norm_added_q = RMSNorm(dim_head, eps)
                self.norm_added_k = RMSNorm(dim_head, eps)
            else:
                self.norm_added_q = self.norm_added_k = None
        out_features = out_dim if out_dim is not None else query_dim
        self.to_out = nn.ModuleList(
            [nn.Linear(inner, out_features, bias=True), nn.Dropout(0.0)])
        if context_pre_only is not None and not context_pre_only:
            self.to_add_out = nn.Linear(inner, out_features, bias=True)

    def _heads(self, t):
        b, s, _ = t.shape
        return t.view(b, s, self.heads, self.dim_head).transpose(1, 2)

    def forward(self, hidden_states, encoder_hidden_states=None,
                attention_mask=None):
        b = hidden_states.shape[0]
        q = self._heads(self.to_q(hidden_states))
        k = self._heads(self.to_k(hidden_states))
        v = self._heads(self.to_v(hidden_states))
        if self.norm_q is not None:
            q = self.norm_q(q)
            k = self.norm_k(k)
        joint = self.added_kv_proj_dim is not None and \
            encoder_hidden_states is not None
        if joint:
            cq = self._heads(self.add_q_proj(encoder_hidden_states))
            ck = self._heads(self.add_k_proj(encoder_hidden_states))
            cv = self._heads(self.add_v_proj(encoder_hidden_states))
            if self.norm_added_q is not None:
                cq = self.norm_added_q(cq)
                ck = self.norm_added_k(ck)
            # token order [sample ; context]
            q = torch.cat([q, cq], dim=2)
            k = torch.cat([k, ck], dim=2)
            v = torch.cat([v, cv], dim=2)
        mask = None
        if attention_mask is not None:
            # [B, q, k] -> repeat_interleave(heads) -> [B, heads, q, k]
            mask = attention_mask
            if mask.shape[0] < b * self.heads:
                mask = mask.repeat_interleave(self.heads, dim=0)
            mask = mask.view(b, self.heads, -1, mask.shape[-1])
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask,
                                           scale=self.scale)
        o = o.transpose(1, 2).reshape(b, -1, self.heads * self.dim_head)
        o = o.to(q.dtype)
        if joint:
            s = hidden_states.shape[1]
            o, co = o[:, :s], o[:, s:]
            o = self.to_out[0](o)
            if not self.context_pre_only:
                co = self.to_add_out(co)
            return o, co
        return self.to_out[0](o)


class JointTransformerBlock(nn.Module):
    """SD3 / SD3.5 MMDiT block, called at crossview_temporal_dit.py:517-521 (A.6)."""

    def __init__(self, dim, num_attention_heads, attention_head_dim,
                 context_pre_only=False, qk_norm=None,
                 use_dual_attention=False):
        super().__init__()
        self.use_dual_attention = use_dual_attention
        self.context_pre_only = context_pre_only
        self.norm1 = SD35AdaLayerNormZeroX(dim) if use_dual_attention \
            else AdaLayerNormZero(dim)
        self.norm1_context = AdaLayerNormContinuous(dim, dim) \
            if context_pre_only else AdaLayerNormZero(dim)
        self.attn = Attention(
            dim, num_attention_heads, attention_head_dim, bias=True,
            qk_norm=qk_norm, eps=1e-6, added_kv_proj_dim=dim,
            context_pre_only=context_pre_only, out_dim=dim)
        if use_dual_attention:
            self.attn2 = Attention(
                dim, num_attention_heads, attention_head_dim, bias=True,
                qk_norm=qk_norm, eps=1e-6, out_dim=dim)
        else:
            self.attn2 = None
        self.norm2 = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.ff = FeedForward(dim, dim, activation_fn="gelu-approximate")
        if not context_pre_only:
            self.norm2_context = nn.LayerNorm(
                dim, elementwise_affine=False, eps=1e-6)
            self.ff_context = FeedForward(
                dim, dim, activation_fn="gelu-approximate")
        else:
            self.norm2_context = None
            self.ff_context = None

    def forward(self, hidden_states, encoder_hidden_states, temb):
        if self.use_dual_attention:
            nh, gate_msa, shift_mlp, scale_mlp, gate_mlp, nh2, gate_msa2 = \
                self.norm1(hidden_states, temb)
        else:
            nh, gate_msa, shift_mlp, scale_mlp, gate_mlp = \
                self.norm1(hidden_states, temb)
        if self.context_pre_only:
            nc = self.norm1_context(encoder_hidden_states, temb)
        else:
            nc, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = \
                self.norm1_context(encoder_hidden_states, temb)

        attn_output, context_attn_output = self.attn(nh, nc)
        hidden_states = hidden_states + gate_msa.unsqueeze(1) * attn_output
        if self.use_dual_attention:
            attn_output2 = self.attn2(nh2)
            hidden_states = hidden_states + \
                gate_msa2.unsqueeze(1) * attn_output2

        nh = self.norm2(hidden_states)
        nh = nh * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
        hidden_states = hidden_states + gate_mlp.unsqueeze(1) * self.ff(nh)

        if self.context_pre_only:
            encoder_hidden_states = None
        else:
            encoder_hidden_states = encoder_hidden_states + \
                c_gate_msa.unsqueeze(1) * context_attn_output
            nc = self.norm2_context(encoder_hidden_states)
            nc = nc * (1 + c_scale_mlp[:, None]) + c_shift_mlp[:, None]
            encoder_hidden_states = encoder_hidden_states + \
                c_gate_mlp.unsqueeze(1) * self.ff_context(nc)
        return encoder_hidden_states, hidden_states


# -- T2I adapter blocks (adapters.py:20) ----------------------------------------

class AdapterResnetBlock(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.block1 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.act = nn.ReLU()
        self.block2 = nn.Conv2d(channels, channels, kernel_size=1)

    def forward(self, x):
        return self.block2(self.act(self.block1(x))) + x


class AdapterBlock(nn.Module):
    def __init__(self, in_channels, out_channels, num_res_blocks, down=False):
        super().__init__()
        self.downsample = nn.AvgPool2d(2, 2, ceil_mode=True) if down else None
        self.in_conv = nn.Conv2d(in_channels, out_channels, 1) \
            if in_channels != out_channels else None
        self.resnets = nn.Sequential(
            *[AdapterResnetBlock(out_channels) for _ in range(num_res_blocks)])

    def forward(self, x):
        if self.downsample is not None:
            x = self.downsample(x)
        if self.in_conv is not None:
            x = self.in_conv(x)
        return self.resnets(x)


# -- scheduler (A.8) -------------------------------------------------------------

class FlowMatchEulerDiscreteSchedulerBase:
    """diffusers FlowMatchEulerDiscreteScheduler: only what
    temporal_independent.py:173-197 and ctsd.py:2024-2026,2056 use."""

    def __init__(self, num_train_timesteps=1000, shift=1.0):
        self.num_train_timesteps = num_train_timesteps
        self.shift = shift
        t = np.linspace(1, num_train_timesteps, num_train_timesteps,
                        dtype=np.float32)[::-1].copy()
        sigmas = torch.from_numpy(t) / num_train_timesteps
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.timesteps = sigmas * num_train_timesteps
        self.sigmas = sigmas
        self.sigma_min = self.sigmas[-1].item()
        self.sigma_max = self.sigmas[0].item()
        self.num_inference_steps = None

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        t = np.linspace(self.sigma_max * self.num_train_timesteps,
                        self.sigma_min * self.num_train_timesteps,
                        num_inference_steps)
        sigmas = t / self.num_train_timesteps
        sigmas = self.shift * sigmas / (1 + (self.shift - 1) * sigmas)
        sigmas = torch.from_numpy(sigmas).to(dtype=torch.float32, device=device)
        self.timesteps = (sigmas * self.num_train_timesteps).to(device=device)
        self.sigmas = torch.cat(
            [sigmas, torch.zeros(1, device=sigmas.device)])
