"""Parity of the B200-native CogVideoX decoder against the fp32 oracle restatement:
single chunk, chunked decode with causal caches (odd and even frame counts), the
diffusion-forcing single-frame decode, and state_dict key parity."""
import pytest
import torch

CFG = dict(block_out_channels=(32, 64, 64, 128), layers_per_block=1, norm_num_groups=8)


def _pair(dtype):
    from oracle import cogvideox as oc
    from dwm.models.cogvideox_vae import AutoencoderKLCogVideoX
    torch.manual_seed(0)
    o = oc.AutoencoderKLCogVideoXDecoder(**CFG)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in o.named_parameters():
            if p.dim() == 1 and "norm_layer.weight" in n:
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
            elif p.dim() == 1:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            else:
                fan = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * fan ** -0.5)
    m = AutoencoderKLCogVideoX(**CFG, compute_dtype=dtype)
    return o, m


def test_state_dict_keys_cpu():
    o, m = _pair(torch.bfloat16)
    so, sm = o.state_dict(), m.state_dict()
    assert set(so) == set(sm), sorted(set(so) ^ set(sm))[:8]
    for k in so:
        assert so[k].shape == sm[k].shape, k
    z = torch.randn(1, 16, 3, 4, 6)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.decode(z)


@pytest.mark.gpu
@pytest.mark.parametrize("frames,dtype,tol", [
    (1, torch.float16, 5e-3), (2, torch.float16, 5e-3), (3, torch.bfloat16, 3e-2),
    (5, torch.float16, 5e-3), (5, torch.bfloat16, 3e-2)])
def test_decode_matches_oracle(frames, dtype, tol):
    o, m = _pair(dtype)
    m.load_state_dict(o.state_dict())
    o, m = o.cuda(), m.cuda()
    g = torch.Generator().manual_seed(frames)
    z = torch.randn(2, 16, frames, 4, 7, generator=g).cuda()
    with torch.no_grad():
        ref = o.decode(z)
    y = m.decode(z, return_dict=False)[0]
    assert y.shape == ref.shape and y.shape[0] == 2 and y.shape[-2:] == (32, 56)
    err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
    assert err < tol, err
