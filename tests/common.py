"""Shared builders for the model-level tests: a tiny CTSD-3.5-shaped config, a seeded
oracle instance (non-zero zero-convs / perturbed mix factors so that adapter and blend
bugs cannot hide), and seeded synthetic conditions."""
import torch

TINY = dict(
    sample_size=16, patch_size=2, in_channels=16, num_layers=4,
    attention_head_dim=64, num_attention_heads=2, joint_attention_dim=64,
    caption_projection_dim=128, pooled_projection_dim=32, out_channels=16,
    pos_embed_max_size=24, dual_attention_layers=[0, 1], qk_norm="rms_norm",
    qk_norm_on_additional_modules="rms_norm", perspective_modeling_type="implicit",
    projection_class_embeddings_input_dim=13 * 256, enable_crossview=True,
    crossview_attention_type="rowwise", crossview_block_layers=[1],
    enable_temporal=True, temporal_attention_type="pointwise",
    temporal_block_layers=[2, 3], merge_factor=2,
    condition_image_adapter_config=dict(
        in_channels=6, channels=[128, 128], is_downblocks=[True, False],
        num_res_blocks=2, downscale_factor=8, use_zero_convs=True))


def seeded_oracle(cfg, seed=0, std=0.05):
    from oracle import ctsd as octsd
    torch.manual_seed(seed)
    m = octsd.DiTCrossviewTemporalConditionModel(**cfg)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.endswith("mix_factor"):
                p.copy_(torch.tensor([0.3]) + 0.5 * torch.randn(1, generator=g))
            elif name.endswith(".weight") and p.dim() == 1:     # norm weights
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith(".bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(std * torch.randn(p.shape, generator=g))
    return m.eval()


def ring_mask(V):
    m = torch.zeros(V, V, dtype=torch.bool)
    for i in range(V):
        for d in (-1, 0, 1):
            m[i, (i + d) % V] = True
    return m


def synthetic_inputs(cfg, B=2, T=4, V=3, H=8, W=12, L=10, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    cond = dict(
        encoder_hidden_states=r(B, T, V, L, cfg["joint_attention_dim"]) * 0.5,
        pooled_projections=r(B, T, V, cfg["pooled_projection_dim"]),
        condition_image_tensor=torch.rand(B, T, V, 6, H * 8, W * 8, generator=g),
        disable_crossview=torch.tensor([False] * B),
        disable_temporal=torch.tensor([False] * B),
        crossview_attention_mask=ring_mask(V).unsqueeze(0).repeat(B, 1, 1),
        added_time_ids=r(B, T, V, 13) * 2,
    )
    sample = r(B, T, V, cfg["in_channels"], H, W)
    timestep = torch.rand(B, T, V, generator=g) * 1000
    mv = lambda t: t.to(device)
    return mv(sample), mv(timestep), {k: mv(v) for k, v in cond.items()}


# -- cases shared by tests/golden/make_reference_golden.py (which runs the REFERENCE's own code
#    on them) and the tests that check the oracle / the CUDA path against those fixtures --------

VARIANTS = ["base", "temporal_rowwise", "temporal_full", "crossview_full", "no_adapter",
            "no_qknorm_extra", "disabled_batch1", "no_perspective", "five_dim"]


def variant_case(name):
    """-> (cfg, sample, timestep, cond, forward kwargs); shared with the test."""
    cfg = dict(TINY)
    extra = {}
    if name == "temporal_rowwise":
        cfg["temporal_attention_type"] = "rowwise"
    elif name == "temporal_full":
        cfg["temporal_attention_type"] = "full"
    elif name == "crossview_full":
        cfg["crossview_attention_type"] = "full"
    elif name == "no_adapter":
        cfg["condition_image_adapter_config"] = None
    elif name == "no_qknorm_extra":
        cfg["qk_norm_on_additional_modules"] = None
    elif name == "no_perspective":
        cfg["perspective_modeling_type"] = ""
    elif name == "five_dim":
        cfg.update(enable_crossview=False, crossview_block_layers=None,
                   perspective_modeling_type="", condition_image_adapter_config=None)
    if name == "five_dim":
        sample, timestep, cond = synthetic_inputs(cfg, V=1)
        sample, timestep = sample.squeeze(2), timestep.squeeze(2)
        cond = dict(encoder_hidden_states=cond["encoder_hidden_states"].squeeze(2),
                    pooled_projections=cond["pooled_projections"].squeeze(2),
                    disable_temporal=cond["disable_temporal"].unsqueeze(1))
        extra["return_dict"] = True
    else:
        sample, timestep, cond = synthetic_inputs(cfg)
        if name == "crossview_full":
            cond["crossview_attention_mask"] = None
        if name == "disabled_batch1":
            cond["disable_temporal"] = torch.tensor([False, True])
            cond["disable_crossview"] = torch.tensor([True, False])
    return cfg, sample, timestep, cond, extra


def scheduler_inputs():
    g = torch.Generator().manual_seed(11)
    shape = (2, 4, 3, 4, 6, 5)
    return dict(
        sample=torch.randn(shape, generator=g), model_output=torch.randn(shape, generator=g),
        noise=torch.randn(shape, generator=g),
        fm_indices=torch.randint(0, 12, shape[:3], generator=g),
        ddim_timesteps=(torch.randint(0, 50, shape[:3], generator=g) * 20 + 1),
        ddpm_timesteps=torch.randint(0, 1000, shape[:3], generator=g))


def _rigid(g, *lead, scale=1.0):
    """Random rigid transforms [..., 4, 4] (rotation about z + translation)."""
    ang = torch.randn(*lead, generator=g) * 0.2
    t = torch.randn(*lead, 3, generator=g) * scale
    m = torch.eye(4).repeat(*lead, 1, 1)
    m[..., 0, 0], m[..., 0, 1] = torch.cos(ang), -torch.sin(ang)
    m[..., 1, 0], m[..., 1, 1] = torch.sin(ang), torch.cos(ang)
    m[..., :3, 3] = t
    return m


def condition_batch(T=4, V=3, L=10, seed=0, hw=(16, 24), text_dim=64, pooled_dim=32):
    """A data batch with every key CrossviewTemporalSD.get_conditions reads (pre-encoded
    text instead of prompts), non-trivial camera / ego matrices."""
    g = torch.Generator().manual_seed(seed)
    K = torch.zeros(1, T, V, 3, 3)
    K[..., 0, 0] = 500 + 50 * torch.rand(1, T, V, generator=g)
    K[..., 1, 1] = 500 + 50 * torch.rand(1, T, V, generator=g)
    K[..., 0, 2], K[..., 1, 2], K[..., 2, 2] = hw[1] / 2, hw[0] / 2, 1.0
    ego = _rigid(g, 1, 1, V + 2, scale=0.5).repeat(1, T, 1, 1, 1)
    step = _rigid(g, 1, T, 1, scale=0.8)
    pose = torch.eye(4).view(1, 1, 1, 4, 4).repeat(1, T, 1, 1, 1)
    for t in range(1, T):
        pose[:, t] = pose[:, t - 1] @ step[:, t]
    ego = pose @ ego                                      # moving rig, [1, T, V + 2, 4, 4]
    ring = torch.zeros(V, V, dtype=torch.bool)
    for i in range(V):
        for d in (-1, 0, 1):
            ring[i, (i + d) % V] = True
    return {
        "pts": torch.arange(T).float().view(1, T, 1).repeat(1, 1, V) * 100,
        "fps": torch.tensor([10.0]),
        "text_embeddings": torch.randn(1, T, V, L, text_dim, generator=g) * 0.5,
        "pooled_text_embeddings": torch.randn(1, T, V, pooled_dim, generator=g),
        "3dbox_images": torch.rand(1, T, V, 3, *hw, generator=g),
        "hdmap_images": torch.rand(1, T, V, 3, *hw, generator=g),
        "crossview_mask": ring.unsqueeze(0),
        "camera_intrinsics": K,
        "camera_transforms": _rigid(g, 1, T, V, scale=1.5),
        "image_size": torch.tensor([float(hw[1]), float(hw[0])]).expand(1, T, V, 2).clone(),
        "ego_transforms": ego,
    }


CONDITION_COMMON = {
    "frame_prediction_style": "diffusion_forcing", "condition_on_all_frames": True,
    "uncondition_image_color": 0.1255, "added_time_ids": "fps_camera_transforms",
    "camera_intrinsic_embedding_indices": [0, 4, 2, 5],
    "camera_intrinsic_denom_embedding_indices": [1, 1, 0, 1],
    "camera_transform_embedding_indices": [2, 6, 10, 3, 7, 11]}
CONDITION_CASES = {
    # name -> (common_config overrides, get_conditions kwargs)
    "cfg": ({}, dict(do_classifier_free_guidance=True)),
    "action_cfg": ({"added_time_ids": "fps_camera_transforms_action",
                    "camera_ego_sensor_indices": [1, 2, 3]},
                   dict(do_classifier_free_guidance=True)),
    "masks_first_frame_only": ({"condition_on_all_frames": False, "disable_temporal": True},
                               dict(do_classifier_free_guidance=False,
                                    _3dbox_condition_mask=torch.tensor([[[True, False, True]]]),
                                    hdmap_condition_mask=torch.tensor([[[False, True, True]]]))),
}


# -- orchestration traces of autoregressive_inference_pipeline ---------------------------------

AUTOREGRESSIVE_CASES = {
    # name -> (common_config, inference_config, latent_shape, total frames, temporal VAE?)
    "windows_ref2": ({"frame_prediction_style": "ctsd"},
                     {"inference_steps": 8, "sequence_length_per_iteration": 6,
                      "reference_frame_count": 2}, (1, 6, 3, 4, 2, 3), 14, False),
    "temporal_vae_ref1": ({"frame_prediction_style": "ctsd"},
                          {"inference_steps": 8, "sequence_length_per_iteration": 17,
                           "reference_frame_count": 1, "vae_pre": 1, "vae_stride": 4,
                           "generate_frames_for_reference": True},
                          (1, 5, 3, 4, 2, 3), 33, True),
    "windows_encoded_reference": ({"frame_prediction_style": "ctsd", "memory_efficient_batch": 4},
                                  {"inference_steps": 8, "sequence_length_per_iteration": 6,
                                   "reference_frame_count": 2,
                                   "generate_frames_for_reference": False},
                                  (1, 6, 3, 4, 2, 3), 14, False),
    "temporal_vae_encoded_reference": ({"frame_prediction_style": "ctsd"},
                                       {"inference_steps": 8, "sequence_length_per_iteration": 17,
                                        "reference_frame_count": 1, "vae_pre": 1, "vae_stride": 4,
                                        "generate_frames_for_reference": False},
                                       (1, 5, 3, 4, 2, 3), 33, True),
    "diffusion_forcing": ({"frame_prediction_style": "diffusion_forcing"},
                          {"inference_steps": 12, "sequence_length_per_iteration": 4,
                           "autoregression_data_exception_for_take_sequence":
                               ["crossview_mask"]},
                          (1, 4, 3, 4, 2, 3), 9, False),
    "diffusion_forcing_clear1": ({"frame_prediction_style": "diffusion_forcing"},
                                 {"inference_steps": 12, "sequence_length_per_iteration": 4,
                                  "clear_reference_frame_count": 1,
                                  "reference_frame_count": 1},
                                 (1, 4, 3, 4, 2, 3), 10, True),
}


class _TraceVae:
    """Stand-in image VAE for the orchestration traces (encode = subsample, no arithmetic of a
    real VAE): `.encode(x).latent_dist.mode()`, `.config`, `.dtype`."""
    class config:
        scaling_factor, shift_factor = 0.5, 0.25
    dtype = torch.float32

    @staticmethod
    def encode(x):
        z = torch.cat([x, x[:, :1]], 1)[..., ::8, ::8]      # image [n,c,h,w] or clip [n,c,t,h,w]
        dist = type("D", (), {"mode": lambda self: z})()
        return type("O", (), {"latent_dist": dist})()


def autoregressive_batch(n_frames, V=3):
    g = torch.Generator().manual_seed(77)
    return {"pts": torch.arange(n_frames).float().view(1, n_frames, 1).repeat(1, 1, V),
            "vae_images": torch.rand(1, n_frames, V, 3, 16, 24, generator=g),
            "fps": torch.tensor([10.0]),
            "crossview_mask": torch.ones(1, V, V, dtype=torch.bool),
            "clip_text": [["frame %d" % t for t in range(n_frames)]]}


def install_fake_inference_pipeline(pipe, trace):
    """Replaces `pipe.inference_pipeline` by a deterministic stand-in that records how the
    orchestration calls it (the same stand-in drives the reference and the mirror)."""
    def fake(latent_shape, batch, output_type, image_latents=None, reference_frame_count=0,
             start_timestep=0, stop_timestep=None, take_time=0):
        k = len(trace)
        g = torch.Generator().manual_seed(1000 + k)
        lat = torch.randn(tuple(latent_shape), generator=g)
        B, T, V = latent_shape[:3]
        n_img = B * (1 if stop_timestep is not None else T) * V
        if pipe.is_temporal_vae and stop_timestep is None:
            n_img = B * ((T - 1) * 4 + 1) * V
        img = torch.arange(n_img).float().view(n_img, 1) + 1000 * k
        trace.append({
            "latent_shape": list(latent_shape), "pts": batch["pts"][0, :, 0].tolist(),
            "clip_text": batch["clip_text"][0], "mask_shape": list(batch["crossview_mask"].shape),
            "output_type": output_type,
            "image_latents": None if image_latents is None else
            [list(image_latents.shape), round(float(image_latents.double().sum()), 4)],
            "reference_frame_count": int(reference_frame_count),
            "start_timestep": int(start_timestep),
            "stop_timestep": None if stop_timestep is None else int(stop_timestep),
            "take_time": int(take_time)})
        return {"images": img, "latents": lat}
    pipe.inference_pipeline = fake


def run_autoregressive_case(cls, name):
    """Drives `cls.autoregressive_inference_pipeline` (reference or mirror class) without
    building a pipeline: only the attributes the orchestration reads are set."""
    common, inf, shape, n_frames, temporal = AUTOREGRESSIVE_CASES[name]
    pipe = object.__new__(cls)
    pipe.common_config, pipe.inference_config, pipe.training_config = common, inf, {}
    pipe.device = torch.device("cpu")
    pipe.generator = torch.Generator().manual_seed(0)
    pipe.is_temporal_vae = temporal
    pipe.test_scheduler = type("S", (), {"init_noise_sigma": 1.0})()
    pipe.vae = _TraceVae()
    pipe.vae.encoder = None            # marks "has an encoder" for the temporal-VAE branch
    pipe.image_processor = type("P", (), {"preprocess": staticmethod(lambda t: 2.0 * t - 1.0)})()
    trace = []
    install_fake_inference_pipeline(pipe, trace)
    out = pipe.autoregressive_inference_pipeline(shape, autoregressive_batch(n_frames), "pt")
    return {"calls": trace, "images_shape": list(out["images"].shape),
            "images_sum": float(out["images"].double().sum())}


# -- orchestration trace of the streaming FIFO (send_frame_condition / receive_frame /
#    fifo_inference_pipeline) ---------------------------------------------------------------------

def run_fifo_case(cls, model, n_frames=7):
    """Drives `cls.fifo_inference_pipeline` (reference or mirror StreamingCrossviewTemporalSD)
    with the real get_conditions (text-less) and a recording stand-in for the denoising loop."""
    pipe = object.__new__(cls)
    pipe.common_config = dict(CONDITION_COMMON, added_time_ids="fps_camera_transforms_action",
                              camera_ego_sensor_indices=[1, 2, 3])
    pipe.inference_config = {
        "guidance_scale": 2.0, "inference_steps": 12, "sequence_length_per_iteration": 4,
        "text_prompt_interval": 1,
        "autoregression_data_exception_for_take_sequence": ["crossview_mask"],
        "autoregression_condition_exception_for_take_sequence": [
            "disable_crossview", "disable_temporal", "crossview_attention_mask",
            "camera_intrinsics_norm", "camera2referego", "encoder_hidden_states"]}
    pipe.device, pipe.model_dtype = torch.device("cpu"), torch.float32
    pipe.generator = torch.Generator().manual_seed(0)
    pipe.model = pipe.model_wrapper = model
    pipe.text_encoders = pipe.tokenizers = pipe.tokenizer = None
    pipe.test_scheduler = type("S", (), {"init_noise_sigma": 1.0,
                                         "set_timesteps": lambda self, n, device=None: None})()
    trace = []

    def summary(d):
        return {k: [list(v.shape), round(float(v.double().sum()), 3)]
                for k, v in sorted(d.items()) if v is not None}

    def fake(latent_shape, start_timestep=0, stop_timestep=None, take_time=0):
        k = len(trace)
        lat = torch.randn(tuple(latent_shape), generator=torch.Generator().manual_seed(500 + k))
        trace.append({"start": int(start_timestep), "stop": int(stop_timestep),
                      "take_time": int(take_time), "conditions": summary(pipe.conditions),
                      "latents_in": round(float(pipe.latents.double().sum()), 3)})
        if stop_timestep >= pipe.inference_config["inference_steps"]:
            pipe.frames.append(lat[:, take_time].flatten(0, 1))
        return lat
    pipe.inference_pipeline = fake
    out = pipe.fifo_inference_pipeline((1, 4, 3, 4, 2, 3), condition_batch(T=n_frames), "pt")
    return {"calls": trace, "images_shape": list(out["images"].shape),
            "images_sum": round(float(out["images"].double().sum()), 3)}


# -- the full-sequence inference_pipeline (reference ctsd.py:1439-1654) ---------------------------

FULL_SEQUENCE_CASES = {
    # name -> (inference_config, reference_frame_count)
    "plain": ({"guidance_scale": 3.0, "inference_steps": 3}, 0),
    "reference_frames": ({"guidance_scale": 3.0, "inference_steps": 3}, 1),
    "no_cfg_partial": ({"inference_steps": 4}, 0),
    # diffusion-forcing branch of the NON-streaming class (what the autoregressive loop calls):
    # queue given as image_latents, partial step range, emitted frame = queue slot take_time
    "df_queue_partial": ({"guidance_scale": 2.0, "inference_steps": 6}, 0),
}


def full_sequence_inputs():
    cfg = dict(TINY, projection_class_embeddings_input_dim=11 * 256)
    batch = condition_batch(T=3, V=3, hw=(64, 96), text_dim=cfg["joint_attention_dim"],
                            pooled_dim=cfg["pooled_projection_dim"])
    common = dict(CONDITION_COMMON, frame_prediction_style="ctsd")
    shape = (1, 3, 3, 16, 8, 12)
    image_latents = torch.randn(shape, generator=torch.Generator().manual_seed(21))
    return cfg, batch, common, shape, image_latents


# -- text conditions through (tiny, randomly initialised) Hugging Face encoders -------------------

def tiny_text_stack():
    """Byte-level CLIP tokenizer built in memory plus seeded tiny CLIP-L / CLIP-G / T5 stand-ins
    and a CLIPTextModel (SD-2.1 path).  No files, no downloads."""
    import transformers

    def bytes_to_unicode():
        bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + \
            list(range(0xAE, 0x100))
        cs, n = bs[:], 0
        for b in range(256):
            if b not in bs:
                bs.append(b)
                cs.append(256 + n)
                n += 1
        return [chr(c) for c in cs]
    chars = sorted(set(bytes_to_unicode()))
    vocab = {}
    for c in chars:
        vocab[c] = len(vocab)
    for c in chars:
        vocab[c + "</w>"] = len(vocab)
    merges = [("t", "h"), ("th", "e</w>"), ("c", "a"), ("ca", "r</w>")]
    for a, b in merges:
        vocab[a + b] = len(vocab)
    vocab["<|startoftext|>"], vocab["<|endoftext|>"] = len(vocab), len(vocab) + 1
    tok = transformers.CLIPTokenizer(vocab=vocab, merges=merges)
    tok.model_max_length = 77

    def clip_cfg(hidden, proj):
        return transformers.CLIPTextConfig(
            vocab_size=len(vocab), hidden_size=hidden, intermediate_size=2 * hidden,
            projection_dim=proj, num_hidden_layers=2, num_attention_heads=2,
            max_position_embeddings=77, bos_token_id=vocab["<|startoftext|>"],
            eos_token_id=vocab["<|endoftext|>"], pad_token_id=vocab["<|endoftext|>"])
    torch.manual_seed(1234)
    clip_l = transformers.CLIPTextModelWithProjection(clip_cfg(32, 24)).eval()
    clip_g = transformers.CLIPTextModelWithProjection(clip_cfg(40, 16)).eval()
    t5 = transformers.T5EncoderModel(transformers.T5Config(
        vocab_size=len(vocab) + 2, d_model=96, d_kv=8, d_ff=64, num_layers=2,
        num_heads=2)).eval()
    clip_sd21 = transformers.CLIPTextModel(clip_cfg(48, 48)).eval()
    return tok, [clip_l, clip_g, t5], clip_sd21


TEXT_PROMPTS_FLAT = ["the car drives at night", "a red truck"]
TEXT_PROMPTS_NESTED = [[["front %d" % t, "left %d" % t, "right %d" % t] for t in range(4)]]
TEXT_CASES = {
    # name -> (is_dit, prompts, text_condition_mask, cfg, drop T5)
    "dit_flat_cfg": (True, TEXT_PROMPTS_FLAT, None, True, False),
    "dit_nested_masked_cfg": (True, TEXT_PROMPTS_NESTED,
                              [[[True, False, True], [True, True, True],
                                [False, False, False], [True, True, False]]], True, False),
    "dit_nested_no_t5": (True, TEXT_PROMPTS_NESTED, None, False, True),
    "unet_nested": (False, TEXT_PROMPTS_NESTED, None, True, False),
}


def text_case_batch(name):
    is_dit, prompts, mask, cfg, drop_t5 = TEXT_CASES[name]
    B = len(prompts)
    batch = {"pts": torch.zeros(B, 4, 3), "fps": torch.tensor([10.0] * B), "clip_text": prompts}
    return batch, (B, 4, 3, 4, 2, 3)


def tensor_fingerprint(t):
    """Compact, order-sensitive signature of a tensor for JSON fixtures."""
    f = t.detach().double().flatten()
    idx = torch.linspace(0, f.numel() - 1, 64).long()
    w = torch.arange(f.numel(), dtype=torch.float64) % 97 + 1
    return {"shape": list(t.shape), "dtype": str(t.dtype), "sum": float(f.sum()),
            "weighted": float((f * w).sum()), "abs": float(f.abs().sum()),
            "samples": [float(v) for v in f[idx]]}


def run_text_case(pipeline_cls, model_classes, name, stack):
    """get_conditions text branch of `pipeline_cls` (reference or mirror) on a TEXT_CASES entry;
    model_classes = (DiT-like class, UNet-like class) whose bare instances satisfy the
    pipeline's isinstance checks."""
    tok, encs, clip21 = stack
    is_dit, prompts, mask, cfg, drop_t5 = TEXT_CASES[name]
    batch, shape = text_case_batch(name)
    model = object.__new__(model_classes[0] if is_dit else model_classes[1])
    te = (encs[:2] + [None] if drop_t5 else encs) if is_dit else clip21
    tk = [tok, tok, tok] if is_dit else tok
    rc = pipeline_cls.get_conditions(model, te, tk, {}, shape, batch, "cpu", torch.float32,
                                     text_condition_mask=mask, do_classifier_free_guidance=cfg)
    return {k: tensor_fingerprint(rc[k]) for k in ("encoder_hidden_states", "pooled_projections")
            if rc.get(k) is not None}


# -- preview_pipeline: latent-shape derivation and dispatch ----------------------------------------

PREVIEW_CASES = {
    # name -> (streaming class?, inference_config, frames in the batch, temporal VAE down blocks)
    "single_window": (False, {"inference_steps": 4}, 5, 4),
    "autoregressive": (False, {"inference_steps": 4, "sequence_length_per_iteration": 6,
                               "reference_frame_count": 2}, 14, 4),
    "temporal_vae": (False, {"inference_steps": 4, "sequence_length_per_iteration": 17,
                             "vae_pre": 1, "vae_stride": 4}, 33, 4),
    "streaming": (True, {"inference_steps": 12, "sequence_length_per_iteration": 4}, 9, 3),
}


def run_preview_case(base_cls, streaming_cls, name):
    streaming, inf, frames, n_down = PREVIEW_CASES[name]
    pipe = object.__new__(streaming_cls if streaming else base_cls)
    pipe.inference_config, pipe.common_config, pipe.training_config = inf, {}, {}
    pipe.should_save = False
    pipe.vae = type("V", (), {"config": type("C", (), {
        "down_block_types": ("D",) * n_down, "latent_channels": 16})()})()
    calls = []

    def rec(kind):
        def f(latent_shape, batch, output_type):
            calls.append([kind, list(latent_shape), list(batch["vae_images"].shape), output_type])
            return {"images": torch.zeros(1)}
        return f
    for kind in ("inference_pipeline", "autoregressive_inference_pipeline",
                 "fifo_inference_pipeline"):
        setattr(pipe, kind, rec(kind))
    batch = {"vae_images": torch.zeros(2, frames, 6, 3, 256, 448), "fps": torch.tensor([10.0])}
    pipe.preview_pipeline(batch, "/nonexistent", 0)
    return calls
