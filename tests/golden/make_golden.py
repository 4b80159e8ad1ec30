"""Generates the committed golden fixtures from the fp32 oracle (run on CPU):

  python tests/golden/make_golden.py

tiny_dit_forward.safetensors : seeded tiny CTSD-3.5-shaped DiT (tests/common.py TINY),
    seeded synthetic inputs -> oracle noise prediction, plus one diffusion-forcing
    denoise step (CFG 2.0) of the latents.
df_schedule.json             : INT diffusion-forcing timestep-index tables and
    in-schedule-range masks (reference ctsd.py:2048-2055, 2083-2088) and the
    FlowMatch-Euler sigma / timestep tables (shift 3.0) for several step counts.

The reference itself cannot be imported here (diffusers==0.31.0 missing), so these pin
the ORACLE, not the reference: parity stays "unpinned" (see oracle/d31.py).
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import TINY, seeded_oracle, synthetic_inputs  # noqa: E402
from oracle import ctsd as octsd  # noqa: E402


def main():
    import safetensors.torch
    torch.set_num_threads(1)          # deterministic reduction order
    o = seeded_oracle(TINY)
    sample, timestep, cond = synthetic_inputs(TINY)
    with torch.no_grad():
        y = o(sample, timestep, **cond)[0][0]
    # one DF denoise step: B=1 latents, CFG-doubled conditions (uncond first)
    sched = octsd.FlowMatchEulerDiscreteScheduler(shift=3.0)
    sched.set_timesteps(12)
    lat = sample[:1].clone()
    new_lat, noise_pred = octsd.df_denoise_step(
        o, sched, lat, cond, i=10, steps_per_inference=3, guidance_scale=2.0)
    safetensors.torch.save_file(
        {"noise_pred_forward": y.contiguous(), "df_step_latents": new_lat.contiguous(),
         "df_step_noise_pred": noise_pred.contiguous()},
        os.path.join(HERE, "tiny_dit_forward.safetensors"))

    tables = {"indices": {}, "in_range": {}, "sigmas": {}, "timesteps": {}}
    for steps, T in ((48, 16), (32, 16), (24, 6), (12, 4)):
        spi = steps // T
        key = "{}x{}".format(steps, T)
        tables["indices"][key] = {
            str(take): [octsd.df_timestep_indices(i, T, spi, take)
                        for i in range(steps + (T - 1) * spi)]
            for take in (0, 1, T - 1)}
        tables["in_range"][key] = [
            [bool(b) for b in octsd.df_in_schedule_range(i, T, spi)]
            for i in range(steps)]
        s = octsd.FlowMatchEulerDiscreteScheduler(shift=3.0)
        s.set_timesteps(steps)
        tables["sigmas"][key] = [float(v) for v in s.sigmas]
        tables["timesteps"][key] = [float(v) for v in s.timesteps]
    with open(os.path.join(HERE, "df_schedule.json"), "w") as f:
        json.dump(tables, f)
    print("wrote fixtures; forward max", float(y.abs().max()))


if __name__ == "__main__":
    main()
