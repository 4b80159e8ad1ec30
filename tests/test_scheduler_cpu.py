"""Integer / index work of the diffusion-forcing scheduler path is bit-exact against the
committed golden tables; the oracle reproduces its own golden fixture."""
import json
import os

import pytest

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _tables():
    with open(os.path.join(HERE, "golden", "df_schedule.json")) as f:
        return json.load(f)


def test_df_index_tables_bit_exact():
    from dwm.schedulers.temporal_independent import df_timestep_indices, df_in_schedule_range
    t = _tables()
    for key, by_take in t["indices"].items():
        steps, T = (int(x) for x in key.split("x"))
        spi = steps // T
        for take, rows in by_take.items():
            for i, row in enumerate(rows):
                assert df_timestep_indices(i, T, spi, int(take)) == row
        for i, row in enumerate(t["in_range"][key]):
            assert df_in_schedule_range(i, T, spi) == row


def test_df_schedule_properties():
    """Size-independent properties: indices are non-increasing over frames, stay inside the
    sigma table, and every frame walks 0..steps-1 exactly once while it is in range."""
    from dwm.schedulers.temporal_independent import df_timestep_indices, df_in_schedule_range
    for steps, T in ((48, 16), (240, 40), (16, 16)):
        spi = steps // T
        seen = [[] for _ in range(T)]
        for i in range(steps + (T - 1) * spi):
            idx = df_timestep_indices(i, T, spi)
            assert all(a >= b for a, b in zip(idx, idx[1:]))
            assert all(0 <= v for v in idx)
            rng = df_in_schedule_range(i, T, spi)
            for j in range(T):
                if rng[j] and idx[j] < steps and i - j * spi < steps:
                    seen[j].append(idx[j])
        for j in range(T):
            assert seen[j] == list(range(steps)), j


def test_flowmatch_tables_match_golden_and_oracle():
    from dwm.schedulers.temporal_independent import FlowMatchEulerDiscreteScheduler
    from oracle import ctsd as octsd
    t = _tables()
    for key in t["sigmas"]:
        steps = int(key.split("x")[0])
        s = FlowMatchEulerDiscreteScheduler(shift=3.0)
        s.set_timesteps(steps)
        assert [float(v) for v in s.sigmas] == t["sigmas"][key]
        assert [float(v) for v in s.timesteps] == t["timesteps"][key]
        o = octsd.FlowMatchEulerDiscreteScheduler(shift=3.0)
        o.set_timesteps(steps)
        assert torch.equal(o.sigmas, s.sigmas)
    s = FlowMatchEulerDiscreteScheduler(shift=3.0)
    s.set_timesteps(48)
    assert s.sigmas.shape == (49,) and s.sigmas[-1] == 0 and s.timesteps[0] == 1000.0
    assert torch.all(s.sigmas[:-1] > s.sigmas[1:])


def test_oracle_reproduces_golden_forward():
    import safetensors.torch
    from common import TINY, seeded_oracle, synthetic_inputs
    g = safetensors.torch.load_file(os.path.join(HERE, "golden", "tiny_dit_forward.safetensors"))
    o = seeded_oracle(TINY)
    sample, timestep, cond = synthetic_inputs(TINY)
    with torch.no_grad():
        y = o(sample, timestep, **cond)[0][0]
    torch.testing.assert_close(y, g["noise_pred_forward"], rtol=1e-4, atol=1e-5)


def test_oracle_known_answers():
    """Closed-form checks of the restated diffusers pieces."""
    import math
    from oracle import d31
    t = torch.tensor([0.0, 1.0, 250.0])
    e = d31.get_timestep_embedding(t, 8, flip_sin_to_cos=True, downscale_freq_shift=0)
    freqs = torch.tensor([math.exp(-math.log(10000) * k / 4) for k in range(4)])
    torch.testing.assert_close(e[:, :4], torch.cos(t[:, None] * freqs))
    torch.testing.assert_close(e[:, 4:], torch.sin(t[:, None] * freqs))
    # AdaLayerNormContinuous chunk order is (scale, shift)
    n = d31.AdaLayerNormContinuous(4, 4)
    with torch.no_grad():
        n.linear.weight.zero_()
        n.linear.bias.copy_(torch.tensor([1., 1, 1, 1, 5, 5, 5, 5]))
    x = torch.randn(2, 3, 4)
    ref = torch.nn.functional.layer_norm(x, (4,), eps=1e-6) * 2 + 5
    torch.testing.assert_close(n(x, torch.zeros(2, 4)), ref)
    # GEGLU: value * gelu(gate), value half first
    g = d31.GEGLU(2, 2)
    with torch.no_grad():
        g.proj.weight.copy_(torch.eye(4)[:, :2] + torch.eye(4)[:, 2:])
        g.proj.bias.zero_()
    v = torch.tensor([[1.0, -2.0]])
    torch.testing.assert_close(g(v), v * torch.nn.functional.gelu(v))
    # RMSNorm
    r = d31.RMSNorm(4, 1e-6)
    x = torch.tensor([[3.0, 4.0, 0.0, 0.0]])
    torch.testing.assert_close(r(x), x / math.sqrt(25 / 4 + 1e-6))
    # PatchEmbed crop is centred
    pe = d31.PatchEmbed(16, 16, 2, 4, 8, 6)
    full = pe.pos_embed.view(6, 6, 8)
    torch.testing.assert_close(pe.cropped_pos_embed(4, 8).view(2, 4, 8), full[2:4, 1:5])


def test_dpm_solver_tables_and_coefficients_match_oracle():
    """`diffusers.DPMSolverMultistepScheduler` mirror (the image example's scheduler): integer
    timesteps and sigma tables equal the oracle restatement; the per-step coefficient table the
    CUDA `step` consumes reproduces the oracle's trajectory when evaluated with plain torch."""
    import torch
    from dwm.schedulers.dpm_solver import DPMSolverMultistepScheduler as M
    from oracle.ctsd import DPMSolverMultistepSchedulerOracle as O
    kw = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", steps_offset=1)
    m = M(prediction_type="v_prediction", **kw)
    m.set_timesteps(50)
    assert m.timesteps[:3].tolist() == [999, 979, 959] and m.timesteps[-1].item() == 20
    assert m.sigmas[-1].item() == 0.0 and m.init_noise_sigma == 1.0
    for pt in ("v_prediction", "epsilon", "sample"):
        for spacing in ("linspace", "leading", "trailing"):
            for n in (50, 10, 1):
                m = M(prediction_type=pt, timestep_spacing=spacing, **kw)
                o = O(prediction_type=pt, timestep_spacing=spacing, **kw)
                m.set_timesteps(n)
                o.set_timesteps(n)
                assert m.timesteps.tolist() == o.timesteps.tolist()     # INT, bit exact
                assert torch.equal(m.sigmas, o.sigmas)
                g = torch.Generator().manual_seed(0)
                x = torch.randn(2, 3, 4, generator=g)
                xo, hist, lower = x.clone(), [None, None], 0
                for i in range(n):
                    v = torch.randn(2, 3, 4, generator=g) * 0.3
                    first = lower < 1 or i == n - 1
                    co = m._coef[i, 0 if first else 1]
                    x0 = co[0] * x + co[1] * v
                    hist = [hist[1], x0]
                    x = co[2] * x + co[3] * x0 + (0 if first else co[4] * hist[0])
                    lower = min(lower + 1, 2)
                    xo = o.step(v, o.timesteps[i], xo)
                assert ((x - xo).abs().max() / xo.abs().max()).item() < 2e-5, (pt, spacing, n)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.step(torch.zeros(2, 3, 4), None, torch.zeros(2, 3, 4))
