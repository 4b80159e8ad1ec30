"""Multi-GPU parity check (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29533 tools/shard_check.py

Every rank builds the same seeded small DiT, runs diffusion-forcing denoise steps with the
CFG x frame sharding plan, and the gathered latents are compared with an unsharded run of
the same steps on rank 0.  Exit code 0 = match."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "src"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from common import TINY, synthetic_inputs
    from dwm.models.crossview_temporal_dit import DiTCrossviewTemporalConditionModel
    from dwm.pipelines.ctsd import StreamingCrossviewTemporalSD
    from opendwm_b200.sharding import ShardPlan
    from opendwm_b200 import lib
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda")
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    V = 3
    use_cfg = os.environ.get("SHARD_CFG", "1") != "0"
    t_ways = world // (2 if (use_cfg and world >= 2) else 1)
    # (frames, temporal attention type): 8 = even shards; 5 / 11 / 19 = the uneven shards of
    # BASELINE configs 5 and 3 (19 with row-wise temporal attention as in config 3)
    cases = [(8, "pointwise"), (5, "pointwise"), (11, "pointwise"), (19, "pointwise"),
             (19, "rowwise"), (5, "rowwise")]
    cases = [c for c in cases if c[0] >= t_ways]
    # row-wise sequences longer than 64 take the tcgen05 kernel when unsharded and the
    # separate-K,V mma.sync kernel when sharded; with attn_tc = 0 both runs use the same kernel
    # and the comparison is bit for bit
    lib.set_option("attn_tc", 0)
    worst, all_equal = 0.0, True
    for T, kind in cases:
        cfg = dict(TINY, temporal_attention_type=kind)
        torch.manual_seed(0)
        model = DiTCrossviewTemporalConditionModel(**cfg, compute_dtype=torch.float16)
        g = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for name, p in model.named_parameters():
                if name.endswith("mix_factor"):
                    p.fill_(0.3)
                elif p.dim() == 1 and name.endswith(".weight"):
                    p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=g))
        steps = 2 * T
        inf = {"guidance_scale": 2.0, "inference_steps": steps,
               "sequence_length_per_iteration": T}
        pipe = StreamingCrossviewTemporalSD(
            None, {"generator_seed": 0}, dev, {"frame_prediction_style": "diffusion_forcing"},
            {}, inf, None, model, model_dtype=torch.float16)
        pipe.reset_streaming((1, T, V, 16, 8, 12), "pt")
        sample, _, cond = synthetic_inputs(cfg, B=2, T=T, V=V, device="cuda")
        cond = {k: (v.half() if v.is_floating_point() and k != "added_time_ids" else v)
                for k, v in cond.items()}
        latents0 = sample[:1].float().contiguous()
        spi = steps // T

        def run(plan):
            pipe.sharding = plan
            pipe.model._cond_key = None
            lat = latents0.clone() if plan is None else plan.local_latents(latents0)
            c = cond if plan is None else plan.local_conditions(cond, cfg_doubled=True)
            fs = slice(0, T) if plan is None else plan.frame_slice()
            for i in (steps - 3, steps - 2, steps - 1):
                idx, ts, in_range = pipe._df_step_tensors(i, T, spi, 0, 1, V)
                pipe.denoise_step(lat, c, idx[:, fs].contiguous(), ts[:, fs].contiguous(),
                                  in_range[fs].contiguous())
            return lat if plan is None else plan.gather_latents(lat)

        plan = ShardPlan(world, rank, T, cfg=use_cfg)
        sharded = run(plan)
        ref = run(None)
        err = ((sharded - ref).abs().max() / ref.abs().max()).item()
        moved = ((ref - latents0).abs().max()).item()
        t = torch.tensor([err, 0.0 if torch.equal(sharded, ref) else 1.0,
                          0.0 if moved > 1e-3 else 1.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst = max(worst, t[0].item())
        all_equal = all_equal and t[1].item() == 0.0 and t[2].item() == 0.0
        if rank == 0:
            print("shard_check world=%d plan=%s T=%d shards=%s temporal=%s peer_scatter=%s "
                  "bit_identical=%s max_rel_err=%.3e (latents moved %.3f)" % (
                      world, plan.parallelism, T, plan.counts, kind,
                      plan.use_peer_scatter and plan.t_ways > 1, t[1].item() == 0.0,
                      t[0].item(), moved), flush=True)
        del pipe, model
        torch.cuda.empty_cache()
    dist.destroy_process_group()
    sys.exit(0 if all_equal and worst == 0.0 else 1)


if __name__ == "__main__":
    main()
