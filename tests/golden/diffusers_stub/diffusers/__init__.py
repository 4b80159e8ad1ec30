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

from torch import nn

from oracle import d31

from . import configuration_utils, image_processor, models, schedulers, utils  # noqa: F401
from .schedulers import DDIMScheduler, DDPMScheduler, FlowMatchEulerDiscreteScheduler  # noqa: F401
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
    """What the reference keeps of diffusers 0.31 UNetSpatioTemporalConditionModel after
    its positional super().__init__ (/root/reference/src/dwm/models/crossview_temporal_unet.py
    :407-418): input / output convolutions and the time / added-time embeddings.  The
    down / mid / up blocks built by the stock class are replaced by the reference at once
    and are therefore not constructed here."""

    @register_to_config
    def __init__(self, sample_size=None, in_channels=8, out_channels=4, down_block_types=(),
                 up_block_types=(), block_out_channels=(320, 640, 1280, 1280),
                 addition_time_embed_dim=256, projection_class_embeddings_input_dim=768,
                 layers_per_block=2, cross_attention_dim=1024, transformer_layers_per_block=1,
                 num_attention_heads=(5, 10, 20, 20), num_frames=25):
        super().__init__()
        boc = block_out_channels
        time_embed_dim = boc[0] * 4
        self.conv_in = nn.Conv2d(in_channels, boc[0], kernel_size=3, padding=1)
        self.time_proj = d31.Timesteps(boc[0], True, 0)
        self.time_embedding = d31.TimestepEmbedding(boc[0], time_embed_dim)
        self.add_time_proj = d31.Timesteps(addition_time_embed_dim, True, 0)
        self.add_embedding = d31.TimestepEmbedding(projection_class_embeddings_input_dim,
                                                   time_embed_dim)
        self.down_blocks = nn.ModuleList([])
        self.up_blocks = nn.ModuleList([])
        self.mid_block = None
        self.conv_norm_out = nn.GroupNorm(num_channels=boc[0], num_groups=32, eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, kernel_size=3, padding=1)


class AutoencoderKL:
    """Name only (default of common_config["vae"], ctsd.py:953-954)."""
