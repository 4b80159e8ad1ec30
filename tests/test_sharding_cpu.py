"""World-size-2 (gloo, CPU) tests of the multi-GPU host logic: the CFG x frame shard
plan, condition slicing, and the index arithmetic of the frame-sharded temporal
attention (local queries against all-gathered K,V), emulated with plain torch."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _run(fn, world):
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_entry, args=(fn, world, port), nprocs=world, join=True)


def _entry(rank, fn, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world)
    finally:
        dist.destroy_process_group()


def _temporal_sharded(rank, world, T=4, kind="pointwise"):
    """Local query frames against the gathered K,V in the unsharded row layout — the index
    arithmetic dwm_b200_attention is given in the frame-sharded temporal call, emulated with
    plain torch for point-wise "(b v hw) t" and row-wise "(b v h) (t w)" regroupings."""
    from opendwm_b200.sharding import ShardPlan
    B, V, Hp, Wp, C = 2, 3, 2, 3, 8
    S = Hp * Wp
    plan = ShardPlan(world, rank, T, cfg=False)          # frames only: t_ways == world
    assert plan.cfg_ways == 1 and plan.t_ways == world and sum(plan.counts) == T
    assert plan.counts[rank] == plan.T_loc and plan.offsets[rank] == plan.t_offset
    g = torch.Generator().manual_seed(0)
    q = torch.randn(B, T, V, S, C, generator=g)
    kv = torch.randn(B, T, V, S, 2 * C, generator=g)
    fs = plan.frame_slice()
    kv_loc = kv[:, fs].reshape(-1, 2 * C).contiguous()
    kv_all = torch.full((B * T * V * S, 2 * C), float("nan"))
    work = plan.gather_frames_kv(kv_loc, kv_all, batch=B, async_op=True)
    work.wait()
    assert torch.equal(kv_all, kv.reshape(-1, 2 * C))     # the unsharded layout, on every rank
    kf, vf = kv_all.view(B, T, V, S, 2 * C).split(C, -1)
    if kind == "pointwise":
        def regroup(t):                                  # "(b t v) hw c -> (b v hw) t c"
            return t.permute(0, 2, 3, 1, 4).reshape(B, V * S, t.shape[1], C)
    else:
        def regroup(t):                                  # "(b t v) (h w) c -> (b v h) (t w) c"
            n = t.shape[1]
            return t.reshape(B, n, V, Hp, Wp, C).permute(0, 2, 3, 1, 4, 5)\
                .reshape(B, V * Hp, n * Wp, C)
    att = torch.softmax(regroup(q[:, fs]) @ regroup(kf).transpose(-1, -2) / C ** 0.5, -1) \
        @ regroup(vf)
    ref = torch.softmax(regroup(q) @ regroup(kv[..., :C]).transpose(-1, -2) / C ** 0.5, -1) \
        @ regroup(kv[..., C:])
    if kind == "pointwise":
        ref = ref[:, :, fs]
    else:
        ref = ref.view(B, V * Hp, T, Wp, C)[:, :, fs].reshape(att.shape)
    torch.testing.assert_close(att, ref)
    # latents round trip over the frame group (uneven shards are padded and trimmed)
    lat = torch.arange(B * T * V, dtype=torch.float32).view(B, T, V)
    assert torch.equal(plan.gather_latents(plan.local_latents(lat)), lat)


def _temporal_sharded_uneven(rank, world):
    _temporal_sharded(rank, world, T=5, kind="pointwise")     # config 5: 5 latent frames
    _temporal_sharded(rank, world, T=5, kind="rowwise")
    _temporal_sharded(rank, world, T=19, kind="rowwise")      # config 3: 19 frames, row-wise
    _temporal_sharded(rank, world, T=11, kind="pointwise")


def _cfg_split(rank, world):
    from opendwm_b200.sharding import ShardPlan
    T, V = 4, 3
    plan = ShardPlan(world, rank, T, cfg=True)
    assert (plan.cfg_ways, plan.t_ways, plan.cfg_rank, plan.T_loc) == (2, 1, rank, 4)
    assert plan.parallelism == "cfg2xframes1"
    cond = {"encoder_hidden_states": torch.arange(2 * T * V * 2.).view(2, T, V, 2),
            "crossview_attention_mask": torch.ones(2, V, V, dtype=torch.bool),
            "disable_temporal": torch.tensor([False, True]), "nothing": None}
    loc = plan.local_conditions(cond, cfg_doubled=True)
    assert loc["encoder_hidden_states"].shape == (1, T, V, 2)
    assert torch.equal(loc["encoder_hidden_states"], cond["encoder_hidden_states"][rank:rank + 1])
    assert loc["disable_temporal"].tolist() == [bool(rank)] and loc["nothing"] is None
    tok = torch.full((6, 4), float(rank))
    both = torch.empty(12, 4)
    plan.gather_cfg_tokens(tok, both)
    assert torch.equal(both[:6], torch.zeros(6, 4)) and torch.equal(both[6:], torch.ones(6, 4))


def test_frame_sharded_temporal_attention_indexing():
    _run(_temporal_sharded, 2)


@pytest.mark.parametrize("world", [2, 4])
def test_uneven_frame_shards_T5_T11_T19(world):
    """5 / 11 / 19 frames over 2 and 4 frame shards (3+2, 2+1+1+1, 5+5+5+4 ...): the gathered
    K,V is the unsharded tensor on every rank, point-wise and row-wise attention of the local
    frames equal the unsharded result, latents round-trip."""
    _run(_temporal_sharded_uneven, world)


def test_cfg_branch_split_and_exchange():
    _run(_cfg_split, 2)


def test_plan_shapes_without_groups():
    from opendwm_b200.sharding import ShardPlan
    seen = set()
    for r in range(8):
        p = ShardPlan(8, r, 16, make_groups=False)
        assert (p.cfg_ways, p.t_ways, p.T_loc) == (2, 4, 4)
        seen.add((p.cfg_rank, p.t_offset))
    assert len(seen) == 8
    with pytest.raises(ValueError):
        ShardPlan(8, 0, 3, make_groups=False)          # fewer frames than frame shards
    # BASELINE configs 5 and 3 on 8 GPUs: uneven frame shards
    for T, want in ((5, [2, 1, 1, 1]), (19, [5, 5, 5, 4]), (11, [3, 3, 3, 2]), (6, [2, 2, 1, 1])):
        plans = [ShardPlan(8, r, T, make_groups=False) for r in range(8)]
        assert plans[0].counts == want
        for c in (0, 1):
            got = [(p.t_offset, p.T_loc) for p in plans if p.cfg_rank == c]
            assert [n for _, n in got] == want
            assert [o for o, _ in got] == [sum(want[:i]) for i in range(4)]
    p1 = ShardPlan(1, 0, 16, make_groups=False)
    assert (p1.cfg_ways, p1.t_ways, p1.T_loc) == (1, 1, 16)


# -- end-to-end sharded window: inference_pipeline with a ShardPlan (slices, per-step index
#    tensors, reference frames crossing shard boundaries, latent gather, item-parallel decode) ----

class _FakeVae:
    class config:
        scaling_factor, shift_factor = 0.5, 0.25
    dtype = torch.float32

    def __init__(self, temporal):
        self.temporal = temporal

    def decode(self, x, return_dict=False):
        if self.temporal:            # [(b v), c, t, h, w] -> 1 + 4 (t - 1) frames
            t_out = 1 + 4 * (x.shape[2] - 1) if x.shape[2] > 1 else 1
            x = x[:, :3].repeat_interleave(4, dim=2)[:, :, :t_out] if x.shape[2] > 1 else x[:, :3]
        else:
            x = x[:, :3]
        return (x * 2 + 1,)


def _window_pipe(plan, df, temporal):
    from dwm.models.crossview_temporal_dit import DiTCrossviewTemporalConditionModel
    from dwm.pipelines.ctsd import CrossviewTemporalSD
    pipe = object.__new__(CrossviewTemporalSD)
    pipe.common_config = {"frame_prediction_style": "diffusion_forcing" if df else "ctsd",
                          "condition_on_all_frames": True, "added_time_ids": "fps_camera_transforms",
                          "camera_intrinsic_embedding_indices": [0, 4, 2, 5],
                          "camera_intrinsic_denom_embedding_indices": [1, 1, 0, 1],
                          "camera_transform_embedding_indices": [2, 6, 10, 3, 7, 11]}
    pipe.inference_config = {"guidance_scale": 2.0, "inference_steps": 8,
                             "sequence_length_per_iteration": 4}
    pipe.device, pipe.model_dtype = torch.device("cpu"), torch.float32
    pipe.generator = torch.Generator().manual_seed(0)
    pipe.model = object.__new__(DiTCrossviewTemporalConditionModel)
    pipe.text_encoders = pipe.tokenizers = None
    pipe.is_dit, pipe.is_temporal_vae, pipe.vae = True, temporal, _FakeVae(temporal)
    pipe.sharding, pipe._step_cache = plan, {}
    pipe.test_scheduler = type("S", (), {
        "timesteps": torch.linspace(1000, 100, 8), "num_inference_steps": 8,
        "init_noise_sigma": 1.0, "set_timesteps": lambda self, n, device=None: None})()

    def fake_step(latents, conditions, idx, timesteps, in_range=None):
        # frame-local update from quantities both CFG branches share
        cam = conditions["camera_transforms"][:latents.shape[0]].sum((-1, -2))
        upd = timesteps.float() / 1000 + 1e-2 * cam + 1e-3 * idx.float()
        new = latents * 0.9 + upd[..., None, None, None]
        if in_range is not None:
            new = torch.where(in_range.bool().view(1, -1, 1, 1, 1, 1), new, latents)
        latents.copy_(new)
        return latents
    pipe.denoise_step = fake_step
    return pipe


def _sharded_window(rank, world, T=4):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from common import condition_batch
    from opendwm_b200.sharding import ShardPlan
    V = 3
    shape = (1, T, V, 4, 2, 3)
    plan = ShardPlan(world, rank, T, cfg=True)
    batch = condition_batch(T=T, V=V)
    ref_lat = torch.randn(shape, generator=torch.Generator().manual_seed(5))
    cases = [dict(df=False, temporal=False, kw={}),
             dict(df=False, temporal=False, kw=dict(image_latents=ref_lat, reference_frame_count=1)),
             dict(df=False, temporal=True, kw=dict(image_latents=ref_lat, reference_frame_count=3,
                                                  start_timestep=2, stop_timestep=7)),
             dict(df=True, temporal=False, kw=dict(image_latents=ref_lat, start_timestep=3,
                                                   stop_timestep=6, take_time=1)),
             dict(df=True, temporal=True, kw={})]
    if T != 4:      # the diffusion-forcing cases need inference_steps % T == 0 (8 steps, T = 4)
        cases = [c for c in cases if not c["df"]]
    for c in cases:
        want = _window_pipe(None, c["df"], c["temporal"]).inference_pipeline(
            shape, batch, "pt", **c["kw"])
        got = _window_pipe(plan, c["df"], c["temporal"]).inference_pipeline(
            shape, batch, "pt", **c["kw"])
        assert got["latents"].shape == want["latents"].shape
        assert torch.equal(got["latents"], want["latents"]), (rank, c)
        assert got["images"].shape == want["images"].shape
        assert torch.equal(got["images"], want["images"]), (rank, c)


def test_sharded_window_matches_unsharded_world2():
    _run(_sharded_window, 2)          # cfg2 x frames1


def test_sharded_window_matches_unsharded_world4():
    _run(_sharded_window, 4)          # cfg2 x frames2, decode of 3 items over 4 ranks


def _seed_broadcast(rank, world):
    """Attaching a plan to a pipeline WITHOUT `generator_seed` makes every rank draw rank 0's
    noise (`generator.seed()` seeds each rank differently and the CFG pair / frame
    shards would silently denoise different latents); with a configured seed nothing changes."""
    from dwm.pipelines.ctsd import CrossviewTemporalSD
    from opendwm_b200.sharding import ShardPlan
    plan = ShardPlan(world, rank, 4, cfg=True)
    pipe = object.__new__(CrossviewTemporalSD)
    pipe.config = {}
    pipe.generator = torch.Generator()
    pipe.generator.manual_seed(1000 + rank)               # stands for generator.seed()
    pipe.sharding = plan
    seeds = [None] * world
    dist.all_gather_object(seeds, pipe.generator.initial_seed())
    assert seeds == [1000] * world, seeds
    noise = torch.randn(3, generator=pipe.generator)
    got = [None] * world
    dist.all_gather_object(got, noise.tolist())
    assert all(g == got[0] for g in got)
    pipe2 = object.__new__(CrossviewTemporalSD)
    pipe2.config = {"generator_seed": 7}
    pipe2.generator = torch.Generator()
    pipe2.generator.manual_seed(7)
    pipe2.sharding = plan
    assert pipe2.generator.initial_seed() == 7
    pipe2.sharding = None                                   # detaching is always allowed
    assert pipe2.sharding is None


def test_sharded_pipeline_broadcasts_the_seed_when_none_is_configured():
    _run(_seed_broadcast, 2)


def _sharded_window_uneven(rank, world):
    _sharded_window(rank, world, T=5)      # frames 3 + 2 over the two frame shards of world 4


def test_sharded_window_with_uneven_frame_shards_world4():
    """BASELINE config 5 has 5 latent frames: cfg2 x frames2 = shards of 3 and 2 frames.  The
    whole window (slices, reference frames crossing the uneven boundary, padded latent gather,
    item-parallel decode) equals the unsharded pipeline."""
    _run(_sharded_window_uneven, 4)
