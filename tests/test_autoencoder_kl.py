"""2-D AutoencoderKL decode: state_dict parity with the oracle
restatement, checkpoint round trip through from_pretrained, deprecated attention key names,
and (GPU) decode parity for SD-3.5-style (16 latent channels, no post_quant_conv) and
SD-2.1-style (4 latent channels + post_quant_conv) configurations."""
import json

import pytest
import torch

SD35 = dict(in_channels=3, out_channels=3, block_out_channels=(32, 64, 128, 128),
            layers_per_block=2, latent_channels=16, norm_num_groups=8,
            scaling_factor=1.5305, shift_factor=0.0609, use_quant_conv=False,
            use_post_quant_conv=False)
SD21 = dict(in_channels=3, out_channels=3, block_out_channels=(64, 128), layers_per_block=1,
            latent_channels=4, norm_num_groups=32, scaling_factor=0.18215)


def _oracle(cfg, seed=0):
    from oracle import autoencoder_kl as oa
    torch.manual_seed(seed)
    o = oa.AutoencoderKL(**cfg)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in o.named_parameters():
            if p.dim() == 1 and "norm" in n and n.endswith(".weight"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
            elif n.endswith(".bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(torch.randn(p.shape, generator=g) * (1.5 / p[0].numel()) ** 0.5)
    return o.eval()


def test_state_dict_and_loader_cpu(tmp_path):
    from dwm.models.autoencoder_kl import AutoencoderKL
    for cfg in (SD35, SD21):
        o = _oracle(cfg)
        m = AutoencoderKL(**cfg)
        so, sm = o.state_dict(), m.state_dict()
        assert set(so) == set(sm), sorted(set(so) ^ set(sm))[:8]
        for k in so:
            assert so[k].shape == sm[k].shape, k
        m.load_state_dict(so, strict=True)                  # full checkpoint
        # a decoder-only checkpoint loads too (encoder keeps its initial weights)
        dec_only = {k: v for k, v in so.items()
                    if not (k.startswith("encoder.") or k.startswith("quant_conv."))}
        m2 = AutoencoderKL(**cfg)
        m2.load_state_dict(dec_only, strict=True)
        assert torch.equal(m2.decoder.conv_in.weight, so["decoder.conv_in.weight"])
    # pre-0.20 attention names with 1x1-conv shaped weights
    old = {}
    for k, v in _oracle(SD21).state_dict().items():
        for new, dep in (("to_q", "query"), ("to_k", "key"), ("to_v", "value"),
                         ("to_out.0", "proj_attn")):
            if ".attentions.0." + new + "." in k:
                k = k.replace(new, dep)
                v = v[:, :, None, None] if v.dim() == 2 else v
        old[k] = v
    m = AutoencoderKL(**SD21)
    m.load_state_dict(old)
    assert torch.equal(m.decoder.mid_block.attentions[0].to_q.weight,
                       _oracle(SD21).decoder.mid_block.attentions[0].to_q.weight)
    # from_pretrained(path, subfolder="vae") round trip
    import safetensors.torch
    d = tmp_path / "vae"
    d.mkdir()
    with open(d / "config.json", "w") as f:
        json.dump(dict(SD35, _class_name="AutoencoderKL", _diffusers_version="0.31.0"), f)
    safetensors.torch.save_file(_oracle(SD35).state_dict(),
                                str(d / "diffusion_pytorch_model.safetensors"))
    v = AutoencoderKL.from_pretrained(str(tmp_path), subfolder="vae")
    assert v.config.scaling_factor == 1.5305 and v.config.shift_factor == 0.0609
    assert not hasattr(v, "post_quant_conv")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        v.decode(torch.zeros(1, 16, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        v.encode(torch.zeros(1, 3, 64, 64))


@pytest.mark.gpu
@pytest.mark.parametrize("name,dtype,tol", [("sd35", torch.float16, 4e-3),
                                            ("sd35", torch.bfloat16, 3e-2),
                                            ("sd21", torch.float16, 4e-3)])
def test_decode_matches_oracle(name, dtype, tol):
    from dwm.models.autoencoder_kl import AutoencoderKL
    cfg = SD35 if name == "sd35" else SD21
    o = _oracle(cfg).cuda()
    m = AutoencoderKL(**cfg, compute_dtype=dtype)
    m.load_state_dict(o.state_dict())
    m.cuda()
    g = torch.Generator().manual_seed(3)
    z = torch.randn(3, cfg["latent_channels"], 8, 12, generator=g).cuda()
    with torch.no_grad():
        ref = o.decode(z, return_dict=False)[0]
    y = m.decode(z.to(dtype), return_dict=False)[0]
    up = 2 ** (len(cfg["block_out_channels"]) - 1)
    assert y.shape == ref.shape == (3, 3, 8 * up, 12 * up) and y.dtype == dtype
    # the decode takes 16-bit latents: compare against the oracle on the same rounded input
    with torch.no_grad():
        ref = o.decode(z.to(dtype).float(), return_dict=False)[0]
    err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
    assert err < tol, err


@pytest.mark.gpu
def test_softmax_rows():
    from opendwm_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(37, 1000, generator=g) * 4).cuda()
    xs = x[:, :968]                                   # row stride != cols
    for dt in (torch.bfloat16, torch.float16):
        out = torch.empty(37, 968, device="cuda", dtype=dt)
        ops.softmax_rows(xs, 0.37, out)
        ref = torch.softmax(xs * 0.37, dim=-1)
        assert (out.float() - ref).abs().max().item() < (4e-3 if dt == torch.bfloat16 else 5e-4)
        assert abs(out.float().sum(-1) - 1).max().item() < 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("name,dtype,tol", [("sd35", torch.float16, 6e-3),
                                            ("sd21", torch.float16, 6e-3)])
def test_encode_matches_oracle(name, dtype, tol):
    from dwm.models.autoencoder_kl import AutoencoderKL
    cfg = SD35 if name == "sd35" else SD21
    o = _oracle(cfg).cuda()
    m = AutoencoderKL(**cfg, compute_dtype=dtype)
    m.load_state_dict(o.state_dict())
    m.cuda()
    g = torch.Generator().manual_seed(5)
    up = 2 ** (len(cfg["block_out_channels"]) - 1)
    x = (torch.rand(3, 3, 8 * up, 12 * up, generator=g) * 2 - 1).cuda()
    with torch.no_grad():
        ref = o.encode(x.to(dtype).float()).latent_dist
    d = m.encode(x.to(dtype)).latent_dist
    assert d.mode().shape == ref.mode().shape == (3, cfg["latent_channels"], 8, 12)
    err = ((d.mode().float() - ref.mode()).abs().max() / ref.mode().abs().max()).item()
    assert err < tol, err
    err = ((d.std.float() - ref.std).abs().max() / ref.std.abs().max()).item()
    assert err < 5 * tol, err
