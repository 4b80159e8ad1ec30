"""Parity at BASELINE.json's FULL north-star size (CTSD-3.5 DFoT 6 views x 16 frames,
CFG-doubled [2,16,6,16,32,56], 24 joint blocks + 18 graft blocks + adapter): one
diffusion-forcing denoise step of the bf16 CUDA path against the fp32 oracle run in eager
PyTorch on the same GPU (TF32 off) with identical weights, latents, indices and conditions.

north_star tolerance: per-step output (the updated latents) max-abs-rel error < 1e-3.
The noise prediction itself (bf16 compute, 42 blocks deep) is also reported and bounded.
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_northstar_step_matches_fp32_oracle():
    import bench
    from oracle import ctsd as octsd
    from dwm.models.crossview_temporal_dit import DiTCrossviewTemporalConditionModel
    from dwm.pipelines.ctsd import StreamingCrossviewTemporalSD
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = bench.load_config()
    B, T, V, C, H, W = cfg["latent_shape"]
    steps = cfg["inference_steps"]
    spi = steps // T
    dev = torch.device("cuda", 0)
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(dev):
            model = DiTCrossviewTemporalConditionModel(**cfg["model"],
                                                       compute_dtype=torch.bfloat16)
    finally:
        torch.set_default_dtype(torch.float32)
    bench.init_weights_(model)
    with torch.no_grad():          # non-trivial blends: sigmoid(mix_factor) away from 0/1
        for n, p in model.named_parameters():
            if n.endswith("mix_factor"):
                p.fill_(0.4)
    with torch.device(dev):
        oracle = octsd.DiTCrossviewTemporalConditionModel(**cfg["model"])
    missing, unexpected = oracle.load_state_dict(model.state_dict(), strict=False)
    assert not missing and not unexpected, (missing[:4], unexpected[:4])
    oracle.eval()

    pipe = StreamingCrossviewTemporalSD(
        None, {"generator_seed": 0}, dev, {"frame_prediction_style": "diffusion_forcing"}, {},
        {"guidance_scale": cfg["guidance_scale"], "inference_steps": steps,
         "sequence_length_per_iteration": T}, None, model, model_dtype=torch.bfloat16)
    pipe.reset_streaming((B, T, V, C, H, W), "pt")
    cond = bench.synthetic_conditions(cfg, 2 * B, T, V, dev, torch.bfloat16)
    lat0 = torch.randn(B, T, V, C, H, W, generator=torch.Generator().manual_seed(0)).to(dev)
    i = steps - 2
    idx, ts, in_range = pipe._df_step_tensors(i, T, spi, 0, B, V)
    lat = lat0.clone()
    pipe.denoise_step(lat, cond, idx, ts, in_range)
    torch.cuda.synchronize()

    sched = octsd.FlowMatchEulerDiscreteScheduler(shift=3.0)
    sched.set_timesteps(steps)
    sched.timesteps = sched.timesteps.to(dev)
    sched.sigmas = sched.sigmas.to(dev)
    cond32 = {k: (v.float() if v.is_floating_point() else v) for k, v in cond.items()}
    with torch.no_grad():
        ref, noise_ref = octsd.df_denoise_step(
            oracle, sched, lat0.clone(), cond32, i=i, steps_per_inference=spi,
            guidance_scale=cfg["guidance_scale"])
    # the native path's noise prediction, recovered from its update: v = (x' - x) / dsigma
    sig = sched.sigmas.to(dev)
    ds = (sig[idx.long() + 1] - sig[idx.long()]).view(B, T, V, 1, 1, 1)
    noise = (lat - lat0) / ds
    err_lat = ((lat - ref).abs().max() / ref.abs().max()).item()
    err_noise = ((noise - noise_ref).abs().max() / noise_ref.abs().max()).item()
    rms_noise = ((noise - noise_ref).pow(2).mean().sqrt() / noise_ref.pow(2).mean().sqrt()).item()
    res = {"shape": [2 * B, T, V, C, H, W], "step_index": i,
           "latents_max_abs_rel": err_lat, "noise_pred_max_abs_rel": err_noise,
           "noise_pred_rms_rel": rms_noise,
           "noise_ref_absmax": noise_ref.abs().max().item(),
           "latents_changed": (lat - lat0).abs().max().item()}
    print(json.dumps(res))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "ns_parity.json"), "w") as f:
            json.dump(res, f)
    assert res["latents_changed"] > 0
    assert err_lat < 1e-3, res
    assert err_noise < 3e-2, res
