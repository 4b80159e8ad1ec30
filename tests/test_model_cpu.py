"""CPU-side checks of the drop-in model mirror: state_dict key/shape parity with the
oracle restatement of the reference module tree, loud failure without CUDA."""
import pytest
import torch

from common import TINY, seeded_oracle, synthetic_inputs


def test_state_dict_keys_match_oracle():
    from dwm.models.crossview_temporal_dit import DiTCrossviewTemporalConditionModel
    o = seeded_oracle(TINY)
    m = DiTCrossviewTemporalConditionModel(**TINY)
    so, sm = o.state_dict(), m.state_dict()
    assert set(so) == set(sm), (sorted(set(so) ^ set(sm))[:10])
    for k in so:
        assert so[k].shape == sm[k].shape, k
    missing, unexpected = m.load_state_dict(so, strict=True)
    assert not missing and not unexpected
    torch.testing.assert_close(m.pos_embed.pos_embed, o.pos_embed.pos_embed,
                               rtol=0, atol=1e-6)


def test_northstar_key_inventory():
    """Key names of SURVEY.md Appendix B exist with the documented shapes (meta device)."""
    from dwm.models.crossview_temporal_dit import DiTCrossviewTemporalConditionModel
    cfg = dict(TINY, num_layers=3, dual_attention_layers=[0], crossview_block_layers=[1],
               temporal_block_layers=[2])
    m = DiTCrossviewTemporalConditionModel(**cfg)
    sd = m.state_dict()
    D = 128
    for k, shape in {
        "pos_embed.proj.weight": (D, 16, 2, 2),
        "time_text_embed.timestep_embedder.linear_1.weight": (D, 256),
        "context_embedder.weight": (D, 64),
        "transformer_blocks.0.norm1.linear.weight": (9 * D, D),
        "transformer_blocks.1.norm1.linear.weight": (6 * D, D),
        "transformer_blocks.2.norm1_context.linear.weight": (2 * D, D),
        "transformer_blocks.0.attn.norm_added_k.weight": (64,),
        "transformer_blocks.0.attn2.to_out.0.bias": (D,),
        "transformer_blocks.0.ff.net.0.proj.weight": (4 * D, D),
        "transformer_blocks.1.ff_context.net.2.weight": (D, 4 * D),
        "norm_out.linear.weight": (2 * D, D), "proj_out.weight": (64, D),
        "condition_image_adapter.body.0.in_conv.weight": (D, 384, 1, 1),
        "condition_image_adapter.body.1.resnets.1.block1.weight": (D, D, 3, 3),
        "condition_image_adapter.zero_convs.0.weight": (D, D, 1, 1),
        "view_embedding.linear_1.weight": (D, 3328),
        "view_pos_embeds.0.linear_1.weight": (4 * D, D),
        "time_pos_embeds.0.linear_2.weight": (D, 4 * D),
        "temporal_transformer_blocks.0.ff_in.net.0.proj.weight": (8 * D, D),
        "crossview_transformer_blocks.0.attn1.to_q.weight": (D, D),
        "crossview_transformer_blocks.0.attn1.norm_q.weight": (64,),
        "view_mixers.0.mix_factor": (1,), "time_mixers.0.mix_factor": (1,),
    }.items():
        assert tuple(sd[k].shape) == shape, (k, tuple(sd[k].shape))
    assert "transformer_blocks.2.attn.to_add_out.weight" not in sd
    assert "crossview_transformer_blocks.0.attn1.to_q.bias" not in sd
    assert "transformer_blocks.1.attn2.to_q.weight" not in sd


def test_no_cpu_fallback():
    from dwm.models.crossview_temporal_dit import DiTCrossviewTemporalConditionModel
    m = DiTCrossviewTemporalConditionModel(**TINY)
    sample, timestep, cond = synthetic_inputs(TINY)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(sample, timestep, **cond)


def test_oracle_runs_and_is_deterministic():
    o = seeded_oracle(TINY)
    sample, timestep, cond = synthetic_inputs(TINY)
    with torch.no_grad():
        y1 = o(sample, timestep, **cond)[0][0]
        y2 = o(sample, timestep, **cond)[0][0]
    assert y1.shape == sample.shape and torch.equal(y1, y2)
    assert torch.isfinite(y1).all() and y1.abs().max() > 1e-3
