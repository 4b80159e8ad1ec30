"""Extra measurements `bench.py` attaches to its N=1 line:

* `streaming_e2e`   — frames/s of the reference-visible streaming unit: one
                      `send_frame_condition(frame data on the host)` + `receive_frame()` with the
                      decoded frame copied back to the host (reference ctsd.py:2105-2232), steady
                      state of the diffusion-forcing FIFO, VAE decode and condition-cache
                      refresh included;
* `workloads`       — the other BASELINE.json configs at full size: config 2 (ctsd_21 6-view image
                      step), config 3 (ctsd_35 6-view x 19-frame row-wise step), config 5
                      (CogVideoX decode of a 6-view x 17-frame window);
* `eager_gpu_baseline` — the oracle restatement of the reference's PyTorch path run eagerly on the
                      SAME GPU with autocast + SDPA (SURVEY.md §2.2 "the real bar"); the oracle is
                      used as the timed baseline only, like `cpu_baseline`.

Model blocks come from tests/golden/example_pipeline_blocks.json (the `pipeline` blocks of the
reference's example configs, extracted by tests/golden/make_example_blocks.py).
"""
import gc
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "src"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _blocks():
    with open(os.path.join(ROOT, "tests", "golden", "example_pipeline_blocks.json")) as f:
        return json.load(f)


def _free():
    gc.collect()
    torch.cuda.empty_cache()


def _events(fn, n, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def _ring(V):
    ring = torch.zeros(V, V, dtype=torch.bool)
    for i in range(V):
        for d in (-1, 0, 1):
            ring[i, (i + d) % V] = True
    return ring


def _rigid(g, *lead, scale=1.0):
    """Random rigid transforms [*lead, 4, 4] (rotation by QR, translation N(0, scale))."""
    a = torch.randn(*lead, 3, 3, generator=g)
    q, _ = torch.linalg.qr(a)
    m = torch.zeros(*lead, 4, 4)
    m[..., :3, :3] = q
    m[..., :3, 3] = torch.randn(*lead, 3, generator=g) * scale
    m[..., 3, 3] = 1.0
    return m


def frame_batch(g, V, hw, L, text_dim, pooled_dim, t):
    """One frame of dataset-shaped condition data on the HOST (what the reference's streaming
    caller hands to send_frame_condition): layout images, camera / ego matrices, pre-encoded
    text."""
    K = torch.zeros(1, 1, V, 3, 3)
    K[..., 0, 0] = K[..., 1, 1] = 500.0
    K[..., 0, 2], K[..., 1, 2], K[..., 2, 2] = hw[1] / 2, hw[0] / 2, 1.0
    pin = (lambda t: t.pin_memory()) if torch.cuda.is_available() else (lambda t: t)
    pose = torch.eye(4).view(1, 1, 1, 4, 4).clone()
    pose[..., 0, 3] = 0.8 * t
    ego = pose @ _rigid(torch.Generator().manual_seed(5), 1, 1, V + 2, scale=0.5)
    return {
        "pts": torch.full((1, 1, V), 100.0 * t), "fps": torch.tensor([10.0]),
        "text_embeddings": pin(torch.randn(1, 1, V, L, text_dim, generator=g) * 0.1),
        "pooled_text_embeddings": torch.randn(1, 1, V, pooled_dim, generator=g),
        "3dbox_images": pin(torch.rand(1, 1, V, 3, *hw, generator=g)),
        "hdmap_images": pin(torch.rand(1, 1, V, 3, *hw, generator=g)),
        "crossview_mask": _ring(V).unsqueeze(0),
        "camera_intrinsics": K,
        "camera_transforms": _rigid(torch.Generator().manual_seed(6), 1, 1, V, scale=1.5),
        "image_size": torch.tensor([float(hw[1]), float(hw[0])]).expand(1, 1, V, 2).clone(),
        "ego_transforms": ego,
    }


def streaming_e2e(model, cfg, dev, dtype, frames=8):
    """Steady-state frames/s of the streaming FIFO through the reference-facing calls."""
    from dwm.models.autoencoder_kl import AutoencoderKL
    from dwm.pipelines.ctsd import StreamingCrossviewTemporalSD
    B, T, V, C, H, W = cfg["latent_shape"]
    steps = cfg["inference_steps"]
    blk = _blocks()["ctsd_35_df16_6views_video_generation_with_layout.json"]["pipeline"]
    common = {k: v for k, v in blk["common_config"].items()
              if k not in ("autocast", "text_encoder_load_args")}
    inf = dict(blk["inference_config"])
    inf.update(guidance_scale=cfg["guidance_scale"], inference_steps=steps,
               sequence_length_per_iteration=T)
    torch.manual_seed(0)
    with torch.device(dev):
        vae = AutoencoderKL(
            block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=C,
            norm_num_groups=32, scaling_factor=1.5305, shift_factor=0.0609,
            use_quant_conv=False, use_post_quant_conv=False, compute_dtype=dtype)
    common["vae_instance"] = vae
    pipe = StreamingCrossviewTemporalSD(None, {"generator_seed": 0}, dev, common, {}, inf, None,
                                        model, model_dtype=dtype)
    m = cfg["model"]
    g = torch.Generator().manual_seed(3)
    mk = lambda t: frame_batch(g, V, (H * 8, W * 8), cfg["text_tokens"],      # noqa: E731
                               m["joint_attention_dim"], m["pooled_projection_dim"], t)
    pipe.reset_streaming((B, T, V, C, H, W), "pt")
    t0 = time.perf_counter()
    for t in range(T):                       # gathering; the T-th call runs the 48-step warm-up
        pipe.send_frame_condition(mk(t))
    first = pipe.receive_frame()
    torch.cuda.synchronize()
    fill_s = time.perf_counter() - t0
    assert first is not None
    lat, h2d, d2h = [], 0, 0
    for k in range(frames + 1):
        fb = mk(T + k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipe.send_frame_condition(fb)
        img = pipe.receive_frame()
        host = img.cpu()                     # device -> host read of the emitted frame (syncs)
        dt = time.perf_counter() - t0
        if k > 0:                            # the first streamed frame warms the ring path
            lat.append(dt)
        h2d = sum(v.numel() * v.element_size() for v in fb.values() if torch.is_tensor(v))
        d2h = host.numel() * host.element_size()
    ok = bool(torch.isfinite(host).all()) and 0.0 <= float(host.min()) and float(host.max()) <= 1.0
    mean = sum(lat) / len(lat)
    spi = steps // T
    res = {"value": 1.0 / mean, "unit": "frames/s", "ms_per_frame": mean * 1e3,
           "ms_per_frame_min": min(lat) * 1e3, "ms_per_frame_max": max(lat) * 1e3,
           "frames": len(lat), "denoise_steps_per_frame": spi,
           "fill_s": fill_s, "h2d_bytes_per_frame": h2d, "d2h_bytes_per_frame": d2h,
           "frame_shape": list(host.shape), "frame_in_unit_range": ok,
           "condition_ring": bool(inf.get("condition_ring", True)),
           "api": "StreamingCrossviewTemporalSD.send_frame_condition(host frame data) + "
                  "receive_frame().cpu(); %d denoise steps + step-invariant cache refresh + "
                  "SD-3.5 AutoencoderKL decode of the 6 views per frame; wall clock around "
                  "each call pair with a device synchronize on both sides" % spi}
    del pipe, vae
    _free()
    return res


# ------------------------------------------------------------------------------- workloads
def _init(model, seed=0):
    import bench
    bench.init_weights_(model, seed)


def workload_unet_step(dev, dtype, steps=10):
    """BASELINE config 2: ctsd_21 6-view image generation, one CFG-doubled UNet step."""
    from unet_bench import MODEL, F_STEP_TFLOP
    from dwm.models.crossview_temporal_unet import UNetCrossviewTemporalConditionModel as U
    from dwm.pipelines.ctsd import CrossviewTemporalSD
    with torch.device(dev):
        m = U(**MODEL, compute_dtype=dtype)
    _init(m)
    pipe = CrossviewTemporalSD(
        None, {"generator_seed": 0}, dev, {"frame_prediction_style": "ctsd"}, {},
        {"guidance_scale": 3, "inference_steps": 50}, None, m, model_dtype=dtype)
    pipe.test_scheduler.set_timesteps(50, dev)
    B, T, V = 1, 1, 6
    gen = torch.Generator().manual_seed(0)
    cond = dict(
        encoder_hidden_states=(torch.randn(2 * B, T, V, 77, 1024, generator=gen) * 0.1)
        .to(dev, dtype),
        condition_image_tensor=None,
        disable_crossview=torch.zeros(2 * B, dtype=torch.bool, device=dev),
        disable_temporal=torch.ones(2 * B, dtype=torch.bool, device=dev),
        crossview_attention_mask=_ring(V).unsqueeze(0).repeat(2 * B, 1, 1).to(dev),
        added_time_ids=torch.randn(2 * B, T, V, 11, generator=gen).to(dev))
    lat = torch.randn(B, T, V, 4, 32, 56, generator=gen).to(dev)
    ts = [pipe.test_scheduler.timesteps[k].to(torch.int32).expand(B, T, V).contiguous()
          for k in range(50)]
    k = [0]

    def go(fn):
        def run():
            fn(lat, cond, None, ts[k[0] % 50], None)
            k[0] += 1
        return run
    ms_eager = _events(go(pipe.denoise_step), steps, warm=3)
    lat.normal_()
    ms = _events(go(pipe.denoise_step_graphed), steps, warm=3)
    res = {"workload": "config 2: ctsd_21 6-view image step [2,1,6,4,32,56], CFG 3, DDIM",
           "ms_per_step": ms, "steps_per_s": 1000.0 / ms, "tflops": F_STEP_TFLOP / ms * 1e3,
           "ms_per_step_without_cuda_graph": ms_eager, "dtype": str(dtype).split(".")[-1],
           "finite": bool(torch.isfinite(lat).all())}
    del pipe, m
    _free()
    return res


def workload_dit_T19_step(dev, dtype, steps=3):
    """BASELINE config 3: ctsd_35 6-view video generation — row-wise cross-view AND row-wise
    temporal attention, 19 latent frames, CFG 4, full-sequence FlowMatch step (no adapter)."""
    from dwm.models.crossview_temporal_dit import DiTCrossviewTemporalConditionModel
    from dwm.pipelines.ctsd import CrossviewTemporalSD
    blk = _blocks()["ctsd_35_6views_video_generation.json"]["pipeline"]
    mcfg = {k: v for k, v in blk["model"].items() if k != "_class_name"}
    B, T, V, C, H, W = 1, 19, 6, 16, 32, 56
    torch.set_default_dtype(dtype)
    with torch.device(dev):
        m = DiTCrossviewTemporalConditionModel(**mcfg, compute_dtype=dtype)
    torch.set_default_dtype(torch.float32)
    _init(m)
    pipe = CrossviewTemporalSD(
        None, {"generator_seed": 0}, dev, {"frame_prediction_style": "ctsd"}, {},
        {"guidance_scale": 4, "inference_steps": 40}, None, m, model_dtype=dtype)
    pipe.test_scheduler.set_timesteps(40, dev)
    g = torch.Generator().manual_seed(0)
    cond = dict(
        encoder_hidden_states=(torch.randn(2 * B, T, V, 154, 4096, generator=g) * 0.1)
        .to(dev, dtype),
        pooled_projections=torch.randn(2 * B, T, V, 2048, generator=g).to(dev, dtype),
        condition_image_tensor=None,
        disable_crossview=torch.zeros(2 * B, dtype=torch.bool, device=dev),
        disable_temporal=torch.zeros(2 * B, dtype=torch.bool, device=dev),
        crossview_attention_mask=_ring(V).unsqueeze(0).repeat(2 * B, 1, 1).to(dev),
        added_time_ids=torch.randn(2 * B, T, V, 11, generator=g).to(dev))
    lat = torch.randn(B, T, V, C, H, W, generator=g).to(dev)
    tt = pipe.test_scheduler.timesteps.to(dev).float()
    k = [0]

    def run():
        i = 20 + k[0] % 3
        idx = torch.full((B, T, V), i, dtype=torch.int32, device=dev)
        pipe.denoise_step(lat, cond, idx, tt[i].expand(B, T, V).contiguous(), None)
        k[0] += 1
    ms = _events(run, steps, warm=2)
    items = 2 * B * T * V
    res = {"workload": "config 3: ctsd_35 6-view x 19-frame step [2,19,6,16,32,56], row-wise "
                       "cross-view + row-wise temporal attention, CFG 4",
           "ms_per_step": ms, "steps_per_s": 1000.0 / ms, "view_frame_items": items,
           "ms_per_item": ms / items, "dtype": str(dtype).split(".")[-1],
           "finite": bool(torch.isfinite(lat).all())}
    del pipe, m
    _free()
    return res


def workload_tvae_window(dev, dtype):
    """BASELINE config 5's decode: CogVideoX temporal VAE, 6 views x 5 latent frames
    [16,5,32,56] -> 17 frames 256x448, `memory_efficient_batch` = 2 views per call."""
    from vae_bench import decoder_flops
    from dwm.models.cogvideox_vae import AutoencoderKLCogVideoX
    torch.manual_seed(0)
    with torch.device(dev):
        vae = AutoencoderKLCogVideoX(compute_dtype=dtype)
    z = torch.randn(6, 16, 5, 32, 56, device=dev)
    out = [None]

    def run():
        out[0] = [vae.decode(z[i:i + 2], return_dict=False)[0] for i in (0, 2, 4)]
    ms = _events(run, 2, warm=1)
    fl = decoder_flops(vae, 5, 32, 56) * 6
    y = out[0][-1]
    res = {"workload": "config 5 decode: CogVideoX temporal VAE, 6 views x [16,5,32,56] -> "
                       "17 frames 256x448, 2 views per call",
           "ms_per_window": ms, "tflops": fl / ms / 1e9, "out_shape_per_call": list(y.shape),
           "dtype": str(dtype).split(".")[-1], "finite": bool(torch.isfinite(y.float()).all())}
    del vae, out, y
    _free()
    return res


def workloads(dev, dtype):
    res = {}
    for name, fn in (("config2_unet_step", workload_unet_step),
                     ("config3_dit_step_T19", workload_dit_T19_step),
                     ("config5_tvae_window", workload_tvae_window)):
        try:
            res[name] = fn(dev, dtype)
        except Exception as e:                       # noqa: BLE001 — never lose the headline
            res[name] = {"error": repr(e)[:300]}
            _free()
    return res


def eager_gpu_baseline(cfg, dev, dtype, steps=2):
    """The oracle restatement of the reference's PyTorch path, eager on the same GPU, autocast
    to `dtype` + SDPA: one north-star diffusion-forcing step.  Baseline only (nothing of the
    product runs here)."""
    import bench
    from oracle import ctsd as octsd
    B, T, V, C, H, W = cfg["latent_shape"]
    n_steps = cfg["inference_steps"]
    spi = n_steps // T
    torch.set_default_dtype(dtype)
    with torch.device(dev):
        oracle = octsd.DiTCrossviewTemporalConditionModel(**cfg["model"])
    torch.set_default_dtype(torch.float32)
    bench.init_weights_(oracle)
    oracle.to(dev).eval()
    cond = bench.synthetic_conditions(cfg, 2 * B, T, V, dev, dtype)
    x = [torch.randn(B, T, V, C, H, W, generator=torch.Generator().manual_seed(0)).to(dev)]
    sched = octsd.FlowMatchEulerDiscreteScheduler(shift=3.0)
    sched.set_timesteps(n_steps)
    sched.timesteps, sched.sigmas = sched.timesteps.to(dev), sched.sigmas.to(dev)
    k = [0]

    def run():
        with torch.no_grad(), torch.autocast("cuda", dtype=dtype):
            x[0] = octsd.df_denoise_step(
                oracle, sched, x[0], cond, i=n_steps - 3 + k[0] % 3, steps_per_inference=spi,
                guidance_scale=cfg["guidance_scale"], model_dtype=dtype)[0].float()
        k[0] += 1
    ms = _events(run, steps, warm=1)
    res = {"value": 1000.0 / ms, "unit": "steps/s", "ms_per_step": ms, "steps": steps,
           "impl": "oracle restatement of the reference's PyTorch modules, eager, autocast %s + "
                   "SDPA, same GPU, same synthetic inputs and weights init" %
                   str(dtype).split(".")[-1],
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    del oracle, cond, x
    _free()
    return res
