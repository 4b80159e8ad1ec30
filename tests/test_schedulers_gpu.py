"""SD-2.1 scheduler mirrors (DDIM eta=0, DDPM add_noise / get_velocity) with per-(b,t,v)
timesteps against the oracle restatement; integer timestep tables bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu
SD21 = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
            beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False,
            steps_offset=1, prediction_type="v_prediction")


@pytest.mark.parametrize("ptype", ["v_prediction", "epsilon", "sample"])
def test_ddim_step(ptype):
    from dwm.schedulers.temporal_independent import DDIMScheduler
    from oracle import ctsd as o
    s = DDIMScheduler(**dict(SD21, prediction_type=ptype))
    r = o.DDIMSchedulerOracle(prediction_type=ptype)
    s.set_timesteps(50, "cuda")
    r.set_timesteps(50)
    assert s.timesteps.cpu().tolist() == r.timesteps.tolist()          # INT, bit exact
    assert s.timesteps[0].item() == 981 and s.timesteps[-1].item() == 1
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 2, 4, 8, 6, generator=g).cuda()
    v = torch.randn(2, 3, 2, 4, 8, 6, generator=g).cuda()
    t = s.timesteps[torch.tensor([[[0, 49], [10, 10], [48, 25]], [[3, 3], [49, 0], [7, 30]]])]
    y = s.step(v, t, x).prev_sample
    ref = r.step(v, t, x)
    torch.testing.assert_close(y, ref, rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(s.alphas_cumprod.cpu(), r.alphas_cumprod)


def test_ddpm_add_noise_and_velocity():
    from dwm.schedulers.temporal_independent import DDPMScheduler
    from oracle import ctsd as o
    s = DDPMScheduler(**{k: SD21[k] for k in ("num_train_timesteps", "beta_start", "beta_end",
                                              "beta_schedule", "prediction_type")})
    r = o.DDPMSchedulerOracle()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 2, 3, 4, 5, 6, generator=g).cuda()
    n = torch.randn(2, 2, 3, 4, 5, 6, generator=g).cuda()
    t = torch.randint(0, 1000, (2, 2, 3), generator=g).cuda()
    torch.testing.assert_close(s.add_noise(x, n, t), r.add_noise(x, n, t), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(s.get_velocity(x, n, t), r.get_velocity(x, n, t), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("ptype", ["v_prediction", "epsilon", "sample"])
def test_dpm_solver_multistep(ptype):
    """DPM-Solver++ (2M, midpoint) mirror: a whole 10-step trajectory on the GPU against the
    oracle restatement of diffusers' DPMSolverMultistepScheduler."""
    from dwm.schedulers.dpm_solver import DPMSolverMultistepScheduler
    from oracle import ctsd as o
    kw = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", steps_offset=1,
              prediction_type=ptype)
    s, r = DPMSolverMultistepScheduler(**kw), o.DPMSolverMultistepSchedulerOracle(**kw)
    s.set_timesteps(10, "cuda")
    r.set_timesteps(10)
    assert s.timesteps.cpu().tolist() == r.timesteps.tolist()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 2, 3, 4, 8, 6, generator=g)
    xs, xr = x.cuda(), x.clone()
    for i in range(10):
        v = torch.randn(1, 2, 3, 4, 8, 6, generator=g) * 0.3
        xs = s.step(v.cuda(), s.timesteps[i], xs).prev_sample
        xr = r.step(v, r.timesteps[i], xr)
        torch.testing.assert_close(xs.cpu(), xr, rtol=5e-5, atol=5e-5)
