"""TEST INFRASTRUCTURE — a name-mapping shim, NOT diffusers.

`diffusers==0.31.0` (pinned by the reference, requirements.txt:6) is not installable in
this image, so the reference's own modules (`/root/reference/src/dwm/...`) cannot be
imported as they are.  This package provides exactly the `diffusers.*` names those modules
touch on the CTSD path and maps each onto the fp32 restatement in `oracle/d31.py`
(`oracle/ctsd.py` for the scheduler tables).  With it on `sys.path`,
`tests/golden/make_reference_golden.py` imports and RUNS the reference's own
`DiTCrossviewTemporalConditionModel`, `VTSelfAttentionBlock`, `AlphaBlender`,
`ImageAdapter` and schedulers from `/root/reference/src` and records their outputs as
golden fixtures.  That pins the part of the oracle that restates OpenDWM's own code
(regroupings, embeddings, mixing, forward order, per-frame scheduler steps); the
diffusers-side arithmetic underneath stays the unpinned restatement.
"""
from types import SimpleNamespace

import torch
from torch import nn

from oracle import d31

from . import configuration_utils, models, schedulers, utils  # noqa: F401
from .configuration_utils import register_to_config


class ModelMixin(nn.Module):
    """`self.config` (attribute access) fed by register_to_config."""

    @property
    def config(self):
        return SimpleNamespace(**self.__dict__.get("_config_dict", {}))

    @property
    def dtype(self):
        return next(self.parameters()).dtype


class SD3Transformer2DModel(ModelMixin):
    """Members of diffusers 0.31 SD3Transformer2DModel, assembled from oracle/d31.py with
    the constructor signature of that release."""

    @register_to_config
    def __init__(self, sample_size=128, patch_size=2, in_channels=16, num_layers=18,
                 attention_head_dim=64, num_attention_heads=18, joint_attention_dim=4096,
                 caption_projection_dim=1152, pooled_projection_dim=2048, out_channels=16,
                 pos_embed_max_size=96, dual_attention_layers=(), qk_norm=None):
        super().__init__()
        self.out_channels = out_channels if out_channels is not None else in_channels
        self.inner_dim = num_attention_heads * attention_head_dim
        self.pos_embed = d31.PatchEmbed(sample_size, sample_size, patch_size, in_channels,
                                        self.inner_dim, pos_embed_max_size)
        self.time_text_embed = d31.CombinedTimestepTextProjEmbeddings(
            self.inner_dim, pooled_projection_dim)
        self.context_embedder = nn.Linear(joint_attention_dim, caption_projection_dim)
        self.transformer_blocks = nn.ModuleList([
            d31.JointTransformerBlock(
                self.inner_dim, num_attention_heads, attention_head_dim,
                context_pre_only=i == num_layers - 1, qk_norm=qk_norm,
                use_dual_attention=i in dual_attention_layers)
            for i in range(num_layers)])
        self.norm_out = d31.AdaLayerNormContinuous(self.inner_dim, self.inner_dim)
        self.proj_out = nn.Linear(self.inner_dim, patch_size * patch_size * self.out_channels)
        self.gradient_checkpointing = False


class UNetSpatioTemporalConditionModel(ModelMixin):
    """Only the name (isinstance checks); the UNet family is pinned through oracle/unet.py."""
