"""Shared builders for the model-level tests: a tiny CTSD-3.5-shaped config, a seeded
oracle instance (non-zero zero-convs / perturbed mix factors so that adapter and blend
bugs cannot hide), and seeded synthetic conditions."""
import torch

TINY = dict(
    sample_size=16, patch_size=2, in_channels=16, num_layers=4,
    attention_head_dim=64, num_attention_heads=2, joint_attention_dim=64,
    caption_projection_dim=128, pooled_projection_dim=32, out_channels=16,
    pos_embed_max_size=24, dual_attention_layers=[0, 1], qk_norm="rms_norm",
    qk_norm_on_additional_modules="rms_norm", perspective_modeling_type="implicit",
    projection_class_embeddings_input_dim=13 * 256, enable_crossview=True,
    crossview_attention_type="rowwise", crossview_block_layers=[1],
    enable_temporal=True, temporal_attention_type="pointwise",
    temporal_block_layers=[2, 3], merge_factor=2,
    condition_image_adapter_config=dict(
        in_channels=6, channels=[128, 128], is_downblocks=[True, False],
        num_res_blocks=2, downscale_factor=8, use_zero_convs=True))


def seeded_oracle(cfg, seed=0, std=0.05):
    from oracle import ctsd as octsd
    torch.manual_seed(seed)
    m = octsd.DiTCrossviewTemporalConditionModel(**cfg)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.endswith("mix_factor"):
                p.copy_(torch.tensor([0.3]) + 0.5 * torch.randn(1, generator=g))
            elif name.endswith(".weight") and p.dim() == 1:     # norm weights
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith(".bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(std * torch.randn(p.shape, generator=g))
    return m.eval()


def ring_mask(V):
    m = torch.zeros(V, V, dtype=torch.bool)
    for i in range(V):
        for d in (-1, 0, 1):
            m[i, (i + d) % V] = True
    return m


def synthetic_inputs(cfg, B=2, T=4, V=3, H=8, W=12, L=10, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    cond = dict(
        encoder_hidden_states=r(B, T, V, L, cfg["joint_attention_dim"]) * 0.5,
        pooled_projections=r(B, T, V, cfg["pooled_projection_dim"]),
        condition_image_tensor=torch.rand(B, T, V, 6, H * 8, W * 8, generator=g),
        disable_crossview=torch.tensor([False] * B),
        disable_temporal=torch.tensor([False] * B),
        crossview_attention_mask=ring_mask(V).unsqueeze(0).repeat(B, 1, 1),
        added_time_ids=r(B, T, V, 13) * 2,
    )
    sample = r(B, T, V, cfg["in_channels"], H, W)
    timestep = torch.rand(B, T, V, generator=g) * 1000
    mv = lambda t: t.to(device)
    return mv(sample), mv(timestep), {k: mv(v) for k, v in cond.items()}
