"""CTSD-2.1 UNet family: state_dict parity with the oracle restatement, the SD-2.1 key
renamer, and (GPU) forward parity for image (T=1) and video (T>1) shapes."""
import pytest
import torch

UCFG = dict(in_channels=4, out_channels=4, block_out_channels=(64, 128, 128, 128),
            num_attention_heads=(1, 2, 2, 2), cross_attention_dim=96,
            projection_class_embeddings_input_dim=11 * 256, layers_per_block=1,
            enable_rowwise_crossview=True, enable_rowwise_temporal=True,
            condition_image_adapter_config=dict(
                in_channels=6, channels=[64, 64, 128, 128, 128],
                is_downblocks=[False, True, True, True, False], num_res_blocks=1,
                downscale_factor=8, use_zero_convs=True))


def _oracle(cfg, seed=0):
    from oracle import unet
    torch.manual_seed(seed)
    o = unet.UNetCrossviewTemporalConditionModel(**cfg)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in o.named_parameters():
            if n.endswith("mix_factor"):
                p.copy_(torch.tensor([0.2]) + 0.3 * torch.randn(1, generator=g))
            elif p.dim() == 1 and n.endswith(".weight"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
            elif n.endswith(".bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(torch.randn(p.shape, generator=g) * p[0].numel() ** -0.5)
    return o.eval()


def _inputs(B, T, V, H=16, W=24, L=7, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    ring = torch.ones(V, V, dtype=torch.bool)
    if V >= 4:
        ring[0, 2] = ring[2, 0] = False
    return r(B, T, V, 4, H, W), torch.rand(B, T, V, generator=g) * 1000, dict(
        encoder_hidden_states=r(B, T, V, L, 96) * 0.5,
        condition_image_tensor=torch.rand(B, T, V, 6, H * 8, W * 8, generator=g),
        disable_crossview=torch.tensor([False] * B),
        disable_temporal=torch.tensor([False] * B),
        crossview_attention_mask=ring.unsqueeze(0).repeat(B, 1, 1),
        added_time_ids=r(B, T, V, 11) * 2)


def test_state_dict_and_renamer_cpu():
    from dwm.models.crossview_temporal_unet import UNetCrossviewTemporalConditionModel as U
    o = _oracle(UCFG)
    m = U(**UCFG)
    so, sm = o.state_dict(), m.state_dict()
    assert set(so) == set(sm), sorted(set(so) ^ set(sm))[:8]
    for k in so:
        assert so[k].shape == sm[k].shape, k
    m.load_state_dict(so, strict=True)
    sd21 = {"down_blocks.0.resnets.1.conv1.weight": 1, "down_blocks.0.attentions.0.norm.weight": 2,
            "mid_block.resnets.0.norm2.bias": 3, "conv_in.weight": 4}
    conv = U.try_to_convert_state_dict(sd21)
    assert set(conv) == {"down_blocks.0.resnets.1.spatial_res_block.conv1.weight",
                         "down_blocks.0.attentions.0.norm.weight",
                         "mid_block.resnets.0.spatial_res_block.norm2.bias", "conv_in.weight"}
    assert U.try_to_convert_state_dict(so).keys() == so.keys()     # already SVD-style
    x, t, c = _inputs(1, 1, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, t, **c)


UNET_CASES = [(2, 1, 3, "image"), (2, 3, 2, "video"), (1, 2, 4, "pointwise"),
              (2, 2, 2, "disabled")]


def unet_case(B, T, V, variant):
    """-> (cfg, sample, timesteps, cond); shared with tests/golden/make_reference_golden.py."""
    cfg = dict(UCFG)
    if variant == "pointwise":
        cfg.update(enable_rowwise_crossview=False, enable_rowwise_temporal=False)
    x, t, c = _inputs(B, T, V)
    if variant == "pointwise":
        c["crossview_attention_mask"] = None
    if variant == "disabled":
        c["disable_temporal"] = torch.tensor([True, True])
        c["disable_crossview"] = torch.tensor([True, False])
    return cfg, x, t, c


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,V,variant", UNET_CASES)
def test_unet_forward_matches_oracle(B, T, V, variant):
    from dwm.models.crossview_temporal_unet import UNetCrossviewTemporalConditionModel as U
    cfg, x, t, c = unet_case(B, T, V, variant)
    o = _oracle(cfg).cuda()
    m = U(**cfg, compute_dtype=torch.float16)
    m.load_state_dict(o.state_dict())
    m.cuda()
    x, t = x.cuda(), t.cuda()
    c = {k: (v.cuda() if v is not None else None) for k, v in c.items()}
    with torch.no_grad():
        ref = o(x, t, **c)[0]
    out, up, down = m(x, t, **c)
    y = out[0]
    assert y.shape == ref.shape
    err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
    assert err < 6e-3, err
