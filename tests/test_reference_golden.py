"""Fixtures produced by RUNNING THE REFERENCE'S OWN CODE (tests/golden/make_reference_golden.py:
/root/reference/src/dwm models + schedulers imported on the diffusers name-mapping shim) pin

  * the oracle restatement (CPU, fp32): bit-exact on the build host, <= 1e-5 elsewhere;
  * the CUDA path (GPU): DiT forward in fp16 against the reference outputs directly, and the
    mirrored per-frame schedulers.

The fixtures do not need /root/reference at test time."""
import os

import pytest
import torch

from common import VARIANTS, scheduler_inputs, seeded_oracle, variant_case

HERE = os.path.dirname(os.path.abspath(__file__))
SD21 = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
            beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False,
            steps_offset=1)


@pytest.fixture(scope="module")
def golden():
    import safetensors.torch
    return safetensors.torch.load_file(os.path.join(HERE, "golden", "reference_outputs.safetensors"))


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()


@pytest.mark.parametrize("name", VARIANTS)
def test_oracle_dit_matches_reference(name, golden):
    torch.set_num_threads(1)
    cfg, sample, timestep, cond, extra = variant_case(name)
    o = seeded_oracle(cfg)
    with torch.no_grad():
        y = o(sample, timestep, **cond, **extra)
    y = y["noise_pred"] if extra else y[0][0]
    ref = golden["dit_" + name]
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


def _unet_cases():
    from test_unet import UNET_CASES
    return UNET_CASES


@pytest.mark.parametrize("B,T,V,variant", _unet_cases())
def test_oracle_unet_matches_reference(B, T, V, variant, golden):
    from test_unet import _oracle, unet_case
    torch.set_num_threads(1)
    cfg, x, t, c = unet_case(B, T, V, variant)
    with torch.no_grad():
        y = _oracle(cfg)(x, t, **c)[0]
    ref = golden["unet_" + variant]
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


def test_oracle_schedulers_match_reference(golden):
    from oracle import ctsd as octsd
    si = scheduler_inputs()
    fm = octsd.FlowMatchEulerDiscreteScheduler(shift=3.0)
    fm.set_timesteps(12)
    assert torch.equal(fm.sigmas, golden["fm_sigmas"])
    assert torch.equal(fm.timesteps, golden["fm_timesteps"])
    y = fm.step_by_indices(si["model_output"], si["fm_indices"], si["sample"])
    assert torch.equal(y, golden["fm_step_by_indices"])
    for pt in ("v_prediction", "epsilon", "sample"):
        ddim = octsd.DDIMSchedulerOracle(prediction_type=pt, beta_start=0.00085, beta_end=0.012)
        ddim.set_timesteps(50)
        y = ddim.step(si["model_output"], si["ddim_timesteps"], si["sample"])
        assert torch.allclose(y, golden["ddim_step_" + pt], rtol=0, atol=2e-6), pt
    assert torch.equal(ddim.timesteps, golden["ddim_timesteps_50"])
    ddpm = octsd.DDPMSchedulerOracle(beta_start=0.00085, beta_end=0.012)
    assert torch.allclose(ddpm.add_noise(si["sample"], si["noise"], si["ddpm_timesteps"]),
                          golden["ddpm_add_noise"], rtol=0, atol=1e-6)
    assert torch.allclose(ddpm.get_velocity(si["sample"], si["noise"], si["ddpm_timesteps"]),
                          golden["ddpm_get_velocity"], rtol=0, atol=1e-6)


def test_df_index_schedule_matches_reference_loop_arithmetic():
    """The diffusion-forcing index expression of the reference loop (ctsd.py:2048-2055 and
    :2083-2088) evaluated literally, against the oracle and the mirrored helpers."""
    from oracle import ctsd as octsd
    from dwm.schedulers import temporal_independent as ti
    for steps, T in ((48, 16), (32, 16), (24, 6), (12, 4)):
        spi = steps // T
        for take in (0, 1):
            for i in range(take * spi, steps):
                want = [min(i - take * spi, max(0, i - j * spi)) for j in range(T)]
                rng = [i - j * spi >= 0 for j in range(T)]
                assert octsd.df_timestep_indices(i, T, spi, take) == want
                assert ti.df_timestep_indices(i, T, spi, take) == want
                assert octsd.df_in_schedule_range(i, T, spi) == rng
                assert ti.df_in_schedule_range(i, T, spi) == rng


@pytest.mark.gpu
@pytest.mark.parametrize("name", VARIANTS)
def test_cuda_dit_matches_reference(name, golden):
    from dwm.models.crossview_temporal_dit import DiTCrossviewTemporalConditionModel
    cfg, sample, timestep, cond, extra = variant_case(name)
    o = seeded_oracle(cfg)
    m = DiTCrossviewTemporalConditionModel(**cfg, compute_dtype=torch.float16)
    m.load_state_dict(o.state_dict())
    m.cuda()
    cond = {k: (v.cuda() if v is not None else None) for k, v in cond.items()}
    y = m(sample.cuda(), timestep.cuda(), **cond, **extra)
    y = y["noise_pred"] if extra else y[0][0]
    assert _rel(y.cpu(), golden["dit_" + name]) < 4e-3


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,V,variant", _unet_cases())
def test_cuda_unet_matches_reference(B, T, V, variant, golden):
    from dwm.models.crossview_temporal_unet import UNetCrossviewTemporalConditionModel as U
    from test_unet import _oracle, unet_case
    cfg, x, t, c = unet_case(B, T, V, variant)
    m = U(**cfg, compute_dtype=torch.float16)
    m.load_state_dict(_oracle(cfg).state_dict())
    m.cuda()
    c = {k: (v.cuda() if v is not None else None) for k, v in c.items()}
    y = m(x.cuda(), t.cuda(), **c)[0][0]
    assert _rel(y.cpu(), golden["unet_" + variant]) < 6e-3


@pytest.mark.gpu
def test_cuda_schedulers_match_reference(golden):
    from dwm.schedulers import temporal_independent as ti
    si = {k: v.cuda() for k, v in scheduler_inputs().items()}
    fm = ti.FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=3.0)
    fm.set_timesteps(12, "cuda")
    assert torch.equal(fm.sigmas.cpu(), golden["fm_sigmas"])
    y = fm.step_by_indices(si["model_output"], si["fm_indices"], si["sample"],
                           return_dict=False)[0]
    torch.testing.assert_close(y.cpu(), golden["fm_step_by_indices"], rtol=1e-5, atol=1e-5)
    for pt in ("v_prediction", "epsilon", "sample"):
        ddim = ti.DDIMScheduler(prediction_type=pt, **SD21)
        ddim.set_timesteps(50, "cuda")
        y = ddim.step(si["model_output"], si["ddim_timesteps"], si["sample"],
                      return_dict=False)[0]
        torch.testing.assert_close(y.cpu(), golden["ddim_step_" + pt], rtol=2e-5, atol=2e-5)
    assert torch.equal(ddim.timesteps.cpu(), golden["ddim_timesteps_50"])
    ddpm = ti.DDPMScheduler(prediction_type="v_prediction", **SD21)
    torch.testing.assert_close(
        ddpm.add_noise(si["sample"], si["noise"], si["ddpm_timesteps"]).cpu(),
        golden["ddpm_add_noise"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(
        ddpm.get_velocity(si["sample"], si["noise"], si["ddpm_timesteps"]).cpu(),
        golden["ddpm_get_velocity"], rtol=1e-5, atol=1e-5)
