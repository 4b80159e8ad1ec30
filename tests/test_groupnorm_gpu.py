"""GroupNorm statistics + fused apply (+SiLU, + SpatialNorm modulation) against
torch.nn.functional.group_norm in fp32, over the channel / group layouts of the CogVideoX
decoder (fast path), the SD-2.1 UNet (10 / 20 / 40 / 60 / 80 channels per group, widths up to
2560) and tiny test models (1-2 channels per group)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nb,T,H,W,C,G", [
    (2, 3, 8, 14, 128, 32),      # fast path (C/4 divides 256)
    (12, 1, 32, 56, 320, 32),    # UNet level 0: 10 channels per group
    (12, 1, 16, 28, 640, 32),
    (12, 1, 8, 14, 1280, 32),
    (12, 1, 4, 7, 2560, 32),     # up-block concat width, 28 pixels
    (3, 1, 8, 14, 1920, 32),
    (3, 1, 5, 9, 960, 32),
    (2, 2, 6, 10, 64, 32),       # 2 channels per group
    (2, 1, 6, 10, 32, 32),       # 1 channel per group
    (1, 1, 64, 112, 512, 32),    # 2-D VAE mid resolution
    (2, 1, 3, 5, 4608, 32),      # C/4 > 1024: last-resort kernel
])
@pytest.mark.parametrize("silu", [True, False])
def test_groupnorm_matches_torch(nb, T, H, W, C, G, silu):
    from opendwm_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    x = (torch.randn(nb, T, H, W, C, generator=g) * 2 + 0.5).cuda()
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).cuda()
    beta = (0.1 * torch.randn(C, generator=g)).cuda()
    sums = ops.groupnorm_stats(x, G)
    xc = x.permute(0, 4, 1, 2, 3).double()                      # [nb, C, T, H, W]
    grp = xc.reshape(nb, G, -1)
    ref_sums = torch.stack([grp.sum(-1), (grp * grp).sum(-1)], -1)
    assert torch.allclose(sums, ref_sums, rtol=2e-5, atol=1e-2), \
        (sums - ref_sums).abs().max().item()
    out = torch.empty(nb, T, H, W, C, device="cuda", dtype=torch.float16)
    ops.spatialnorm_silu(x, sums, gamma, beta, out, groups=G, eps=1e-6, silu=silu)
    ref = torch.nn.functional.group_norm(xc.float(), G, gamma, beta, eps=1e-6)
    if silu:
        ref = torch.nn.functional.silu(ref)
    ref = ref.permute(0, 2, 3, 4, 1)
    assert (out.float() - ref).abs().max().item() < 4e-3
