"""ORACLE (test infrastructure, never shipped or measured as the product).

fp32 PyTorch restatement of OpenDWM's own CTSD code on the denoising hot path,
each function citing the reference file:line it follows.  Built on oracle/d31.py
for the diffusers pieces.

PINNED AGAINST THE REFERENCE'S OWN CODE for everything restated in THIS file:
tests/golden/make_reference_golden.py imports /root/reference/src/dwm
(crossview_temporal_dit.py, crossview_temporal.py, adapters.py,
schedulers/temporal_independent.py) on a name-mapping shim that resolves `diffusers.*` to
oracle/d31.py, runs nine DiT configurations, the three schedulers and three iterations of
the reference's StreamingCrossviewTemporalSD.inference_pipeline loop, and
tests/test_reference_golden.py checks this module against those outputs (bit-exact on the
build host).  The diffusers-side arithmetic underneath (oracle/d31.py) stays PARITY
UNPINNED (see its header).
"""
import torch
from torch import nn

from . import d31


class AlphaBlender(nn.Module):
    """src/dwm/models/crossview_temporal.py:9-72."""

    def __init__(self, alpha, merge_strategy="learned_with_images"):
        super().__init__()
        self.merge_strategy = merge_strategy
        if merge_strategy == "fixed":
            self.register_buffer("mix_factor", torch.Tensor([alpha]))
        elif merge_strategy in ("learned", "learned_with_images"):
            self.register_parameter(
                "mix_factor", nn.Parameter(torch.Tensor([alpha])))
        else:
            raise ValueError(merge_strategy)

    def get_alpha(self, image_only_indicator=None):
        if self.merge_strategy == "fixed":
            return self.mix_factor
        if self.merge_strategy == "learned":
            return torch.sigmoid(self.mix_factor)
        if image_only_indicator is None:
            raise ValueError("Please provide image_only_indicator")
        return torch.where(
            image_only_indicator,
            torch.ones((1,), device=image_only_indicator.device),
            torch.sigmoid(self.mix_factor))

    def forward(self, a, b, image_only_indicator=None):
        alpha = self.get_alpha(image_only_indicator).to(a.dtype)
        alpha = alpha.view(
            *alpha.shape, *[1 for _ in range(a.dim() - alpha.dim())])
        return alpha * a + (1.0 - alpha) * b


class VTSelfAttentionBlock(nn.Module):
    """src/dwm/models/crossview_temporal.py:536-582."""

    def __init__(self, dim, time_mix_inner_dim, num_attention_heads,
                 attention_head_dim, qk_norm=None):
        super().__init__()
        self.norm_in = nn.LayerNorm(dim)
        self.ff_in = d31.FeedForward(
            dim, dim_out=time_mix_inner_dim, activation_fn="geglu")
        self.norm1 = nn.LayerNorm(time_mix_inner_dim)
        self.attn1 = d31.Attention(
            time_mix_inner_dim, num_attention_heads, attention_head_dim,
            bias=False, qk_norm=qk_norm, eps=1e-5)
        self.norm3 = nn.LayerNorm(time_mix_inner_dim)
        self.ff = d31.FeedForward(time_mix_inner_dim, activation_fn="geglu")

    def forward(self, hidden_states, self_attention_mask=None):
        residual = hidden_states
        hidden_states = self.ff_in(self.norm_in(hidden_states)) + residual
        attn = self.attn1(self.norm1(hidden_states),
                          attention_mask=self_attention_mask)
        hidden_states = attn + hidden_states
        return self.ff(self.norm3(hidden_states)) + hidden_states


class ImageAdapter(nn.Module):
    """src/dwm/models/adapters.py:6-60 (zero_gate_coef path omitted: no shipped
    config on this path sets it)."""

    def __init__(self, in_channels=3, channels=(320, 320, 640, 1280, 1280),
                 is_downblocks=(False, True, True, True, False),
                 num_res_blocks=2, downscale_factor=8, use_zero_convs=False,
                 zero_gate_coef=None, gradient_checkpointing=True):
        super().__init__()
        assert not zero_gate_coef
        in_channels = in_channels * downscale_factor ** 2
        self.unshuffle = nn.PixelUnshuffle(downscale_factor)
        self.body = nn.ModuleList([
            d31.AdapterBlock(in_channels if i == 0 else channels[i - 1],
                             channels[i], num_res_blocks, down=is_downblocks[i])
            for i in range(len(channels))])
        self.zero_convs = nn.ModuleList(
            [nn.Conv2d(c, c, 1) for c in channels]) if use_zero_convs else \
            [None for _ in channels]
        for i in self.zero_convs:
            if i is not None:
                nn.init.zeros_(i.weight)
                nn.init.zeros_(i.bias)

    def forward(self, x):
        base_shape = x.shape[:-3]
        x = self.unshuffle(x.flatten(0, -4))
        features = []
        for block, zero_conv in zip(self.body, self.zero_convs):
            x = block(x)
            x_out = x if zero_conv is None else zero_conv(x)
            features.append(x_out.view(*base_shape, *x_out.shape[1:]))
        return features


class DiTCrossviewTemporalConditionModel(nn.Module):
    """src/dwm/models/crossview_temporal_dit.py:105-630 on top of the
    SD3Transformer2DModel members it inherits (Appendix A/B).  Only the branches
    reachable from dwm.pipelines.ctsd inference are restated: cross-view
    "rowwise"/"full", temporal "pointwise"(else-branch)/"rowwise"/"full",
    perspective_modeling_type "implicit" or "", AlphaBlender mixers."""

    def __init__(self, sample_size=128, patch_size=2, in_channels=16,
                 num_layers=18, attention_head_dim=64, num_attention_heads=18,
                 joint_attention_dim=4096, caption_projection_dim=1152,
                 pooled_projection_dim=2048, out_channels=16,
                 pos_embed_max_size=96, dual_attention_layers=(),
                 qk_norm=None, projection_class_embeddings_input_dim=None,
                 condition_image_adapter_config=None, enable_crossview=False,
                 enable_temporal=False, crossview_attention_type=None,
                 temporal_attention_type=None, merge_factor=2,
                 merge_strategy="learned_with_images",
                 crossview_block_layers=None, temporal_block_layers=None,
                 crossview_gradient_checkpointing=False,
                 temporal_gradient_checkpointing=False,
                 mixer_type="AlphaBlender", perspective_modeling_type="",
                 disable_view_emb_on_temporal_module=False,
                 qk_norm_on_additional_modules=None, mask_module=None):
        super().__init__()
        assert mixer_type == "AlphaBlender" and mask_module is None
        self.patch_size = patch_size
        self.out_channels = out_channels
        inner_dim = attention_head_dim * num_attention_heads
        self.inner_dim = inner_dim
        self.pos_embed = d31.PatchEmbed(
            sample_size, sample_size, patch_size, in_channels, inner_dim,
            pos_embed_max_size)
        self.time_text_embed = d31.CombinedTimestepTextProjEmbeddings(
            inner_dim, pooled_projection_dim)
        self.context_embedder = nn.Linear(
            joint_attention_dim, caption_projection_dim)
        self.transformer_blocks = nn.ModuleList([
            d31.JointTransformerBlock(
                inner_dim, num_attention_heads, attention_head_dim,
                context_pre_only=i == num_layers - 1, qk_norm=qk_norm,
                use_dual_attention=i in dual_attention_layers)
            for i in range(num_layers)])
        self.norm_out = d31.AdaLayerNormContinuous(inner_dim, inner_dim)
        self.proj_out = nn.Linear(
            inner_dim, patch_size * patch_size * out_channels)

        self.disable_view_emb_on_temporal_module = \
            disable_view_emb_on_temporal_module
        self.condition_image_adapter = None \
            if condition_image_adapter_config is None else \
            ImageAdapter(**condition_image_adapter_config)
        self.index_proj = d31.Timesteps(inner_dim, True, 0)
        self.perspective_modeling_type = perspective_modeling_type
        if perspective_modeling_type == "implicit":
            self.view_cam_proj = d31.Timesteps(256, True, 0)
            self.view_embedding = d31.TimestepEmbedding(
                projection_class_embeddings_input_dim, inner_dim)
        elif perspective_modeling_type != "":
            raise NotImplementedError(perspective_modeling_type)

        self.enable_crossview = enable_crossview
        self.crossview_attention_type = crossview_attention_type
        self.crossview_block_layers = crossview_block_layers
        if enable_crossview:
            n = len(crossview_block_layers)
            self.view_pos_embeds = nn.ModuleList([
                d31.TimestepEmbedding(inner_dim, inner_dim * 4, inner_dim)
                for _ in range(n)])
            self.crossview_transformer_blocks = nn.ModuleList([
                VTSelfAttentionBlock(
                    inner_dim, inner_dim, num_attention_heads,
                    attention_head_dim, qk_norm_on_additional_modules)
                for _ in range(n)])
            self.view_mixers = nn.ModuleList([
                AlphaBlender(merge_factor, merge_strategy) for _ in range(n)])
        self.enable_temporal = enable_temporal
        self.temporal_attention_type = temporal_attention_type
        self.temporal_block_layers = temporal_block_layers
        if enable_temporal:
            n = len(temporal_block_layers)
            self.time_pos_embeds = nn.ModuleList([
                d31.TimestepEmbedding(inner_dim, inner_dim * 4, inner_dim)
                for _ in range(n)])
            self.temporal_transformer_blocks = nn.ModuleList([
                VTSelfAttentionBlock(
                    inner_dim, inner_dim, num_attention_heads,
                    attention_head_dim, qk_norm_on_additional_modules)
                for _ in range(n)])
            self.time_mixers = nn.ModuleList([
                AlphaBlender(merge_factor, merge_strategy) for _ in range(n)])

    # crossview_temporal_dit.py:223-327
    def forward_crossview_block_and_mix_result(
            self, block, mixer, hidden_states, view_emb, batch_size,
            sequence_length, view_count, width, height, disable_crossview,
            crossview_attention_mask):
        import einops
        h = hidden_states + view_emb
        if self.crossview_attention_type == "full":
            h = einops.rearrange(h, "(bt v) (h w) c -> bt (h v w) c",
                                 v=view_count, w=width)
            h = block(h, self_attention_mask=crossview_attention_mask)
            h = einops.rearrange(h, "bt (h v w) c -> (bt v) (h w) c",
                                 v=view_count, w=width)
        elif self.crossview_attention_type == "rowwise":
            if crossview_attention_mask is not None:
                crossview_attention_mask = crossview_attention_mask\
                    .repeat_interleave(width, 2)\
                    .repeat_interleave(width, 1)\
                    .repeat_interleave(sequence_length * height, 0)
            h = einops.rearrange(h, "(bt v) (h w) c -> (bt h) (v w) c",
                                 w=width, v=view_count)
            h = block(h, self_attention_mask=crossview_attention_mask)
            h = einops.rearrange(h, "(bt h) (v w) c -> (bt v) (h w) c",
                                 bt=batch_size * sequence_length, v=view_count)
        else:
            raise TypeError("Not support {}".format(
                self.crossview_attention_type))
        return mixer(
            hidden_states.view(batch_size, sequence_length * view_count,
                               *hidden_states.shape[1:]),
            h.reshape(batch_size, sequence_length * view_count, *h.shape[1:]),
            image_only_indicator=disable_crossview).flatten(0, 1)

    # crossview_temporal_dit.py:329-370
    def forward_temporal_block_and_mix_result(
            self, block, mixer, hidden_states, sequence_emb, batch_size,
            sequence_length, view_count, width, disable_temporal):
        import einops
        h = hidden_states + sequence_emb
        if self.temporal_attention_type == "full":
            h = einops.rearrange(h, "(b t v) hw c -> (b v) (t hw) c",
                                 b=batch_size, t=sequence_length)
            h = block(h)
            h = einops.rearrange(h, "(b v) (t hw) c -> (b t v) hw c",
                                 b=batch_size, t=sequence_length)
        elif self.temporal_attention_type == "rowwise":
            h = einops.rearrange(h, "(b t v) (h w) c -> (b v h) (t w) c",
                                 b=batch_size, v=view_count, w=width)
            h = block(h)
            h = einops.rearrange(h, "(b v h) (t w) c -> (b t v) (h w) c",
                                 b=batch_size, v=view_count, w=width)
        else:
            h = einops.rearrange(h, "(b t v) hw c -> (b v hw) t c",
                                 b=batch_size, t=sequence_length)
            h = block(h)
            h = einops.rearrange(h, "(b v hw) t c -> (b t v) hw c",
                                 b=batch_size, v=view_count, t=sequence_length)
        return mixer(
            hidden_states.view(batch_size, sequence_length * view_count,
                               *hidden_states.shape[1:]),
            h.reshape(batch_size, sequence_length * view_count, *h.shape[1:]),
            image_only_indicator=disable_temporal).flatten(0, 1)

    # crossview_temporal_dit.py:372-630
    def forward(self, sample, timestep=None, frustum_bev_residuals=None,
                encoder_hidden_states=None, pooled_projections=None,
                condition_image_tensor=None, disable_crossview=None,
                disable_temporal=None, crossview_attention_mask=None,
                crossview_attention_index=None, camera_intrinsics=None,
                camera_transforms=None, camera_intrinsics_norm=None,
                camera2referego=None, added_time_ids=None, noise=None,
                return_dict=False):
        should_add_dim = sample.dim() < 6
        if should_add_dim:
            sample = sample.unsqueeze(2)
            timestep = timestep.unsqueeze(2)
            if condition_image_tensor is not None:
                condition_image_tensor = condition_image_tensor.unsqueeze(2)
            if encoder_hidden_states is not None:
                encoder_hidden_states = encoder_hidden_states.unsqueeze(2)
            if disable_temporal is not None:
                disable_temporal = disable_temporal.unsqueeze(2)
            if pooled_projections is not None:
                pooled_projections = pooled_projections.unsqueeze(2)

        hidden_states = sample
        batch_size, sequence_length, view_count, _, height, width = \
            hidden_states.shape
        p = self.patch_size
        height, width = height // p, width // p

        hidden_states = hidden_states.flatten(0, 2)
        pooled_projections = pooled_projections.flatten(0, 2)
        encoder_hidden_states = encoder_hidden_states.flatten(0, 2)

        hidden_states = self.pos_embed(hidden_states)
        encoder_hidden_states = self.context_embedder(encoder_hidden_states)
        temb = self.time_text_embed(timestep.flatten(), pooled_projections)

        view_cam_emb = 0
        if self.perspective_modeling_type == "implicit":
            view_emb = self.view_cam_proj(added_time_ids.flatten())\
                .to(dtype=hidden_states.dtype)
            view_cam_emb = self.view_embedding(view_emb.view(
                batch_size * sequence_length * view_count, -1)).unsqueeze(1)

        condition_residuals = None
        if self.condition_image_adapter is not None and \
                condition_image_tensor is not None:
            condition_residuals = self.condition_image_adapter(
                condition_image_tensor)

        for i, block in enumerate(self.transformer_blocks):
            if condition_residuals is not None and len(condition_residuals) > 0:
                hidden_states = hidden_states + \
                    condition_residuals.pop(0).flatten(0, 2)\
                    .flatten(2).permute(0, 2, 1)

            encoder_hidden_states, hidden_states = block(
                hidden_states, encoder_hidden_states, temb)

            if self.enable_temporal and i in self.temporal_block_layers:
                k = self.temporal_block_layers.index(i)
                sequence_emb = torch.arange(
                    sequence_length, device=hidden_states.device)\
                    .unsqueeze(0).unsqueeze(-1).repeat(batch_size, 1, view_count)
                sequence_emb = self.index_proj(sequence_emb.flatten())\
                    .to(dtype=hidden_states.dtype)
                sequence_emb = self.time_pos_embeds[k](sequence_emb).unsqueeze(1)
                if self.enable_crossview and \
                        not self.disable_view_emb_on_temporal_module:
                    sequence_emb = sequence_emb + view_cam_emb
                hidden_states = self.forward_temporal_block_and_mix_result(
                    self.temporal_transformer_blocks[k], self.time_mixers[k],
                    hidden_states, sequence_emb, batch_size, sequence_length,
                    view_count, width, disable_temporal)

            if self.enable_crossview and i in self.crossview_block_layers:
                k = self.crossview_block_layers.index(i)
                view_emb = torch.arange(
                    view_count, device=hidden_states.device)\
                    .unsqueeze(0).unsqueeze(0)\
                    .repeat(batch_size, sequence_length, 1)
                view_emb = self.index_proj(view_emb.flatten())\
                    .to(dtype=hidden_states.dtype)
                view_emb = self.view_pos_embeds[k](view_emb).unsqueeze(1)
                view_emb = view_emb + view_cam_emb
                hidden_states = self.forward_crossview_block_and_mix_result(
                    self.crossview_transformer_blocks[k], self.view_mixers[k],
                    hidden_states, view_emb, batch_size, sequence_length,
                    view_count, width, height, disable_crossview,
                    crossview_attention_mask)

        hidden_states = self.norm_out(hidden_states, temb)
        hidden_states = self.proj_out(hidden_states)
        hidden_states = hidden_states.reshape(
            hidden_states.shape[0], height, width, p, p, self.out_channels)
        hidden_states = torch.einsum("nhwpqc->nchpwq", hidden_states)
        output = hidden_states.reshape(
            batch_size, sequence_length, view_count, self.out_channels,
            height * p, width * p)
        result = [output]
        if return_dict:
            return {"noise_pred": output.squeeze(2) if should_add_dim else output}
        return result, sequence_length, sequence_length


# -- schedulers ------------------------------------------------------------------

class FlowMatchEulerDiscreteScheduler(d31.FlowMatchEulerDiscreteSchedulerBase):
    """src/dwm/schedulers/temporal_independent.py:173-197."""

    def step_by_indices(self, model_output, timestep_indices, sample):
        if isinstance(timestep_indices, torch.Tensor):
            while timestep_indices.dim() < model_output.dim():
                timestep_indices = timestep_indices.unsqueeze(-1)
        sample = sample.to(torch.float32)
        idx = timestep_indices.long().to(self.sigmas.device)
        sigma = self.sigmas[idx]
        sigma_next = self.sigmas[idx + 1]
        prev_sample = sample + (sigma_next - sigma) * model_output
        return prev_sample.to(model_output.dtype)


def df_timestep_indices(i, sequence_length, steps_per_inference, take_time=0):
    """Diffusion-forcing index schedule, ctsd.py:2048-2055 (INT, bit-exact)."""
    return [min(i - take_time * steps_per_inference,
                max(0, i - j * steps_per_inference))
            for j in range(sequence_length)]


def df_in_schedule_range(i, sequence_length, steps_per_inference):
    """ctsd.py:2083-2088."""
    return [i - j * steps_per_inference >= 0 for j in range(sequence_length)]


@torch.no_grad()
def df_denoise_step(model, scheduler, latents, conditions, i, steps_per_inference,
                    guidance_scale=None, take_time=0, model_dtype=torch.float32):
    """One iteration of StreamingCrossviewTemporalSD.inference_pipeline's loop,
    ctsd.py:2046-2090."""
    B, T, V = latents.shape[:3]
    idx = torch.tensor(df_timestep_indices(i, T, steps_per_inference, take_time),
                       dtype=torch.int32, device=latents.device)\
        .unsqueeze(0).unsqueeze(-1).repeat(B, 1, V)
    timesteps = scheduler.timesteps[idx.long()]
    x = latents.to(dtype=model_dtype)
    if guidance_scale is not None:
        x = torch.cat([x, x])
        timesteps_input = torch.cat([timesteps, timesteps])
    else:
        timesteps_input = timesteps
    out, _, _ = model(x, timesteps_input, **conditions)
    noise_pred = out[0]
    if guidance_scale is not None:
        u, c = noise_pred.chunk(2)
        noise_pred = u + guidance_scale * (c - u)
    staging = scheduler.step_by_indices(noise_pred, idx.cpu(), latents)
    in_range = torch.tensor(
        df_in_schedule_range(i, T, steps_per_inference), device=latents.device)\
        .view(1, T, 1, 1, 1, 1)
    return torch.where(in_range, staging, latents), noise_pred


# -- SD-2.1 schedulers (temporal_independent.py:8-170; diffusers DDIM/DDPM tables, A.8) -----

def scaled_linear_alphas(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                           dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class DDPMSchedulerOracle:
    """temporal_independent.py:8-45."""

    def __init__(self, **kw):
        self.alphas_cumprod = scaled_linear_alphas(**kw)

    def add_noise(self, original_samples, noise, timesteps):
        while timesteps.dim() < original_samples.dim():
            timesteps = timesteps.unsqueeze(-1)
        a = self.alphas_cumprod.to(original_samples)[timesteps.long()]
        return a ** 0.5 * original_samples + (1 - a) ** 0.5 * noise

    def get_velocity(self, sample, noise, timesteps):
        while timesteps.dim() < sample.dim():
            timesteps = timesteps.unsqueeze(-1)
        a = self.alphas_cumprod.to(sample)[timesteps.long()]
        return a ** 0.5 * noise - (1 - a) ** 0.5 * sample


class DDIMSchedulerOracle:
    """temporal_independent.py:48-170 with eta = 0 (leading spacing, steps_offset)."""

    def __init__(self, num_train_timesteps=1000, prediction_type="v_prediction",
                 steps_offset=1, set_alpha_to_one=False, **kw):
        self.num_train_timesteps = num_train_timesteps
        self.prediction_type = prediction_type
        self.steps_offset = steps_offset
        self.alphas_cumprod = scaled_linear_alphas(num_train_timesteps, **kw)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one \
            else self.alphas_cumprod[0]

    def set_timesteps(self, n):
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        self.timesteps = (torch.arange(0, n) * ratio).flip(0) + self.steps_offset

    def step(self, model_output, timestep, sample):
        while timestep.dim() < sample.dim():
            timestep = timestep.unsqueeze(-1)
        prev = timestep - self.num_train_timesteps // self.num_inference_steps
        ac = self.alphas_cumprod.to(sample.device)
        a_t = ac[timestep.long()]
        a_p = torch.where(prev >= 0, ac[prev.clamp_min(0).long()],
                          self.final_alpha_cumprod.to(sample.device))
        b_t = 1 - a_t
        if self.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        elif self.prediction_type == "sample":
            x0 = model_output
            eps = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
        else:
            x0 = a_t ** 0.5 * sample - b_t ** 0.5 * model_output
            eps = a_t ** 0.5 * model_output + b_t ** 0.5 * sample
        return a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps


class DPMSolverMultistepSchedulerOracle:
    """diffusers==0.31.0 DPMSolverMultistepScheduler, the scheduler the reference's image
    example selects (examples/ctsd_21_6views_image_generation.json `inference_config.scheduler`,
    consumed at ctsd.py:981-985 and stepped with a scalar timestep at :1573-1575), restated for
    algorithm_type "dpmsolver++", solver_type "midpoint", solver_order <= 2, no Karras sigmas,
    no thresholding (the diffusers defaults on top of the SD-2.1 scheduler_config.json).
    PARITY UNPINNED (restated from the published source)."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02,
                 beta_schedule="linear", solver_order=2, prediction_type="epsilon",
                 lower_order_final=True, euler_at_final=False, final_sigmas_type="zero",
                 timestep_spacing="linspace", steps_offset=0, **unused):
        import numpy as np
        self.np = np
        self.num_train_timesteps = num_train_timesteps
        self.solver_order, self.prediction_type = solver_order, prediction_type
        self.lower_order_final, self.euler_at_final = lower_order_final, euler_at_final
        self.final_sigmas_type = final_sigmas_type
        self.timestep_spacing, self.steps_offset = timestep_spacing, steps_offset
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                   dtype=torch.float32) ** 2
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.init_noise_sigma = 1.0

    def set_timesteps(self, n, device=None):
        np = self.np
        last = self.num_train_timesteps            # lambda_min_clipped = -inf: nothing clipped
        if self.timestep_spacing == "linspace":
            ts = np.linspace(0, last - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif self.timestep_spacing == "leading":
            ratio = last // (n + 1)
            ts = (np.arange(0, n + 1) * ratio).round()[::-1][:-1].copy().astype(np.int64)
            ts += self.steps_offset
        elif self.timestep_spacing == "trailing":
            ratio = self.num_train_timesteps / n
            ts = np.arange(last, 0, -ratio).round().copy().astype(np.int64) - 1
        else:
            raise ValueError(self.timestep_spacing)
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        if self.final_sigmas_type == "zero":
            last_sigma = 0.0
        else:
            last_sigma = float(((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]) ** 0.5)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [last_sigma]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        self.num_inference_steps = n
        self.model_outputs = [None] * self.solver_order
        self.lower_order_nums = 0
        self.step_index = 0

    @staticmethod
    def _alpha_sigma(sigma):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha_t, sigma * alpha_t

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample):
        i = self.step_index
        n = len(self.timesteps)
        lower_final = i == n - 1 and (self.euler_at_final or
                                      (self.lower_order_final and n < 15) or
                                      self.final_sigmas_type == "zero")
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[i])
        if self.prediction_type == "epsilon":
            x0 = (sample - sigma_t * model_output) / alpha_t
        elif self.prediction_type == "sample":
            x0 = model_output
        else:
            x0 = alpha_t * sample - sigma_t * model_output
        for k in range(self.solver_order - 1):
            self.model_outputs[k] = self.model_outputs[k + 1]
        self.model_outputs[-1] = x0
        sample = sample.to(torch.float32)

        def lam(sigma):
            a, s = self._alpha_sigma(sigma)
            return a, s, torch.log(a) - torch.log(s)
        a_t, s_t, l_t = lam(self.sigmas[i + 1])
        a_0, s_0, l_0 = lam(self.sigmas[i])
        h = l_t - l_0
        if self.solver_order == 1 or self.lower_order_nums < 1 or lower_final:
            prev = (s_t / s_0) * sample - (a_t * (torch.exp(-h) - 1.0)) * x0
        else:
            a_1, s_1, l_1 = lam(self.sigmas[i - 1])
            m0, m1 = self.model_outputs[-1], self.model_outputs[-2]
            r0 = (l_0 - l_1) / h
            d0, d1 = m0, (1.0 / r0) * (m0 - m1)
            prev = (s_t / s_0) * sample - (a_t * (torch.exp(-h) - 1.0)) * d0 \
                - 0.5 * (a_t * (torch.exp(-h) - 1.0)) * d1
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev.to(model_output.dtype)
