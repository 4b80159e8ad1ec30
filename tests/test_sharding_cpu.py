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


def _gathered_rows(B, T_loc, V, S, t_ways):
    """Row of key position j = r*T_loc + t in kv_all.view(-1, C) for group (b, v*S+s),
    exactly the arithmetic dwm_b200_attention is given in the sharded temporal call."""
    rows = B * T_loc * V * S
    b, r_, j = torch.meshgrid(torch.arange(B), torch.arange(V * S),
                              torch.arange(t_ways * T_loc), indexing="ij")
    base = b * (T_loc * V * S) + r_
    return base + (j // T_loc) * rows + (j % T_loc) * (V * S)


def _temporal_sharded(rank, world):
    from opendwm_b200.sharding import ShardPlan
    B, T, V, S, C = 2, 4, 3, 5, 8
    plan = ShardPlan(world, rank, T, cfg=False)          # frames only: t_ways == world
    assert (plan.cfg_ways, plan.t_ways, plan.T_loc) == (1, 2, 2)
    g = torch.Generator().manual_seed(0)
    q = torch.randn(B, T, V, S, C, generator=g)
    kv = torch.randn(B, T, V, S, 2 * C, generator=g)
    fs = plan.frame_slice()
    kv_loc = kv[:, fs].reshape(-1, 2 * C).contiguous()
    kv_all = torch.empty(world * kv_loc.shape[0], 2 * C)
    plan.gather_frames_kv(kv_loc, kv_all)
    rows = _gathered_rows(B, plan.T_loc, V, S, plan.t_ways)      # [B, V*S, T]
    k_all = kv_all[rows.reshape(-1), :C].view(B, V * S, T, C)
    v_all = kv_all[rows.reshape(-1), C:].view(B, V * S, T, C)
    q_loc = q[:, fs].permute(0, 2, 3, 1, 4).reshape(B, V * S, plan.T_loc, C)
    att = torch.softmax(q_loc @ k_all.transpose(-1, -2) / C ** 0.5, -1) @ v_all
    # unsharded reference: "(b t v) hw c -> (b v hw) t c"
    qf = q.permute(0, 2, 3, 1, 4).reshape(B, V * S, T, C)
    kf = kv[..., :C].permute(0, 2, 3, 1, 4).reshape(B, V * S, T, C)
    vf = kv[..., C:].permute(0, 2, 3, 1, 4).reshape(B, V * S, T, C)
    ref = torch.softmax(qf @ kf.transpose(-1, -2) / C ** 0.5, -1) @ vf
    torch.testing.assert_close(att, ref[:, :, fs])
    # latents round trip over the frame group
    lat = torch.arange(B * T * V, dtype=torch.float32).view(B, T, V)
    assert torch.equal(plan.gather_latents(plan.local_latents(lat)), lat)


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
        ShardPlan(8, 0, 6, make_groups=False)
    p1 = ShardPlan(1, 0, 16, make_groups=False)
    assert (p1.cfg_ways, p1.t_ways, p1.T_loc) == (1, 1, 16)
