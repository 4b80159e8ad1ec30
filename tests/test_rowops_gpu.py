"""Parity of the row / elementwise kernels against plain fp32 PyTorch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _r(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


@pytest.mark.parametrize("D", [1536, 320, 640, 2048])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_layernorm_affine_add(D, dtype):
    from opendwm_b200 import ops
    items, S = 5, 37
    M = items * S
    x = _r((M, D), 1) * 3 + 0.5
    emb = _r((items, D), 2)
    full = _r((M, D), 3)
    w, b = _r((D,), 4) * 0.1 + 1, _r((D,), 5) * 0.1
    out = torch.empty(M, D, dtype=dtype, device="cuda")
    ssum = torch.empty(M, D, device="cuda")
    ops.layernorm(x, out, weight=w, bias=b, eps=1e-5, add_item=emb, add_full=full,
                  rows_per_item=S, sum_out=ssum)
    t = x + emb.repeat_interleave(S, 0) + full
    ref = torch.nn.functional.layer_norm(t, (D,), w, b, 1e-5)
    assert torch.equal(ssum, t)
    tol = 6e-3 if dtype == torch.bfloat16 else 1e-3
    assert ((out.float() - ref).abs().max() / ref.abs().max()).item() < tol


def test_layernorm_dual_modulation():
    from opendwm_b200 import ops
    items, S, D = 4, 50, 1536
    M = items * S
    x = _r((M, D), 1)
    mod = _r((items, 9 * D), 2) * 0.3
    sh, sc, sh2, sc2 = mod[:, :D], mod[:, D:2 * D], mod[:, 6 * D:7 * D], mod[:, 7 * D:8 * D]
    o1 = torch.empty(M, D, dtype=torch.bfloat16, device="cuda")
    o2 = torch.empty_like(o1)
    ops.layernorm(x, o1, eps=1e-6, rows_per_item=S, shift=sh, scale=sc, shift2=sh2, scale2=sc2, out2=o2)
    n = torch.nn.functional.layer_norm(x, (D,), None, None, 1e-6)
    r1 = n * (1 + sc.repeat_interleave(S, 0)) + sh.repeat_interleave(S, 0)
    r2 = n * (1 + sc2.repeat_interleave(S, 0)) + sh2.repeat_interleave(S, 0)
    assert ((o1.float() - r1).abs().max() / r1.abs().max()).item() < 6e-3
    assert ((o2.float() - r2).abs().max() / r2.abs().max()).item() < 6e-3


@pytest.mark.parametrize("D,items,S,variant", [
    (1536, 13, 448, "mod"),          # exact vector count, item changes inside row groups
    (1536, 11, 397, "dual"),         # ragged last group (M % 8 != 0)
    (320, 41, 101, "affine_item"),   # D/4 = 80: bounds-checked vectors, small rows
    (1280, 9, 500, "affine_item"),
    (2048, 5, 1000, "mod"),
])
@pytest.mark.parametrize("staged", [1, 0])
def test_layernorm_large_m_staged_and_resident(D, items, S, variant, staged):
    """M >= 4096 rows take the bulk-copy staged kernel (ln_staged = 1); both kernels must give
    bit-identical results (same per-row arithmetic) and match torch."""
    from opendwm_b200 import ops, lib
    M = items * S
    xbuf = _r((M, D + 64), 1) * 2 + 0.3
    x = xbuf[:, :D]                                            # row pitch != D
    out = torch.empty(M, D, dtype=torch.bfloat16, device="cuda")
    lib.set_option("ln_staged", staged)
    try:
        if variant == "affine_item":
            emb = _r((items, D), 2)
            w, b = _r((D,), 4) * 0.1 + 1, _r((D,), 5) * 0.1
            ssum = torch.empty(M, D, device="cuda")
            ops.layernorm(x, out, weight=w, bias=b, eps=1e-5, add_item=emb, rows_per_item=S,
                          sum_out=ssum)
            t = x + emb.repeat_interleave(S, 0)
            assert torch.equal(ssum, t)
            ref = torch.nn.functional.layer_norm(t, (D,), w, b, 1e-5)
            outs, refs = [out], [ref]
        else:
            mod = _r((items, 9 * D), 2) * 0.3
            sh, sc = mod[:, :D], mod[:, D:2 * D]
            n = torch.nn.functional.layer_norm(x, (D,), None, None, 1e-6)
            refs = [n * (1 + sc.repeat_interleave(S, 0)) + sh.repeat_interleave(S, 0)]
            outs = [out]
            if variant == "dual":
                sh2, sc2 = mod[:, 6 * D:7 * D], mod[:, 7 * D:8 * D]
                o2 = torch.empty_like(out)
                ops.layernorm(x, out, eps=1e-6, rows_per_item=S, shift=sh, scale=sc, shift2=sh2,
                              scale2=sc2, out2=o2)
                outs.append(o2)
                refs.append(n * (1 + sc2.repeat_interleave(S, 0)) + sh2.repeat_interleave(S, 0))
            else:
                ops.layernorm(x, out, eps=1e-6, rows_per_item=S, shift=sh, scale=sc)
    finally:
        lib.set_option("ln_staged", 1)
    for o, r in zip(outs, refs):
        assert ((o.float() - r).abs().max() / r.abs().max()).item() < 6e-3


def test_act_cast_and_sinusoid():
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from opendwm_b200 import ops, lib
    from oracle import d31
    x = _r((7, 1536), 1)
    o = torch.empty(7, 1536, dtype=torch.bfloat16, device="cuda")
    ops.act_cast(x, o, lib.ACT_SILU)
    assert torch.equal(o, torch.nn.functional.silu(x).bfloat16()) or \
        (o.float() - torch.nn.functional.silu(x)).abs().max() < 2e-2
    t = torch.tensor([0.0, 1.0, 15.0, 999.0, 637.25, -1000.0], device="cuda")
    for ch, flip, shift in [(256, True, 0.0), (1536, True, 0.0), (320, True, 0.0), (256, False, 1.0)]:
        out = torch.empty(t.numel(), ch, dtype=torch.float16, device="cuda")
        ops.sinusoid(t, ch, out, flip, shift)
        ref = d31.get_timestep_embedding(t, ch, flip, shift)
        assert (out.float() - ref).abs().max().item() < 2e-3


def test_patchify_matches_conv():
    from opendwm_b200 import ops, lib
    items, C, H, W, P, D = 3, 16, 8, 12, 2, 256
    x = _r((items, C, H, W), 1)
    conv = torch.nn.Conv2d(C, D, P, P).cuda()
    a = torch.empty(items * (H // P) * (W // P), C * P * P, dtype=torch.float16, device="cuda")
    ops.patchify(x, P, a)
    y = ops.linear(a, conv.weight.detach().reshape(D, -1).half().contiguous(),
                   conv.bias.detach().float(), epilogue=lib.EPI_F32)
    ref = conv(x).flatten(2).transpose(1, 2).reshape(-1, D)
    assert ((y - ref).abs().max() / ref.abs().max()).item() < 2e-3


@pytest.mark.parametrize("cfg", [1, 2])
@pytest.mark.parametrize("rd", [torch.float32, torch.float16])
def test_cfg_euler_step(cfg, rd):
    from opendwm_b200 import ops
    B, T, V, C, H, W, P = 1, 4, 2, 16, 4, 6, 2
    S = (H // P) * (W // P)
    tok = _r((cfg * B * T * V * S, P * P * C), 1)
    lat = _r((B, T, V, C, H, W), 2)
    sig = torch.linspace(1, 0, 13, device="cuda")
    idx = torch.tensor([[[3, 3], [2, 2], [0, 0], [11, 11]]], dtype=torch.int32, device="cuda")
    in_range = torch.tensor([1, 1, 0, 1], dtype=torch.uint8, device="cuda")
    # reference un-patchify (crossview_temporal_dit.py:603-621)
    hs = tok.view(cfg * B * T * V, H // P, W // P, P, P, C)
    npd = torch.einsum("nhwpqc->nchpwq", hs).reshape(cfg * B, T, V, C, H, W)
    if cfg == 2:
        u, c = npd.chunk(2)
        npd = u + 2.5 * (c - u)
    d = (sig[idx.long() + 1] - sig[idx.long()])[..., None, None, None]
    staged = (lat + d * npd).to(rd).float()
    ref = torch.where(in_range.bool().view(1, T, 1, 1, 1, 1), staged, lat)
    got = lat.clone()
    npo = torch.empty_like(lat)
    ops.cfg_euler_step(tok, got, idx, sig, cfg=cfg, guidance_scale=2.5, patch=P,
                       in_range=in_range, noise_pred=npo, round_dtype=rd)
    torch.testing.assert_close(npo, npd, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(got, ref, rtol=1e-6, atol=1e-6 if rd == torch.float32 else 1e-3)
    assert torch.equal(got[:, 2], lat[:, 2])
