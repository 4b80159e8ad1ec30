"""Algorithmic FLOPs (2*MAC) of one CTSD DiT noise-predict forward.

Counts the linears and attentions that depend on latents/timestep; the ImageAdapter,
context_embedder and pooled-text / index embeddings are step-invariant and excluded
(the reference recomputes them every step: 447.4 vs 396.2 TFLOP at the north star)."""


def dit_flops_per_item(D=1536, S=448, L=154, num_layers=24, n_dual=13,
                       n_crossview=6, n_temporal=12, seq_crossview=168,
                       seq_temporal=16, patch_k=64, out_n=64, include_context_embed=True,
                       joint_attention_dim=4096):
    f = 0.0
    for i in range(num_layers):
        last = i == num_layers - 1
        dual = i < n_dual
        f += (24 + (8 if dual else 0)) * S * D * D           # sample linears
        f += (6 if last else 24) * L * D * D                  # context linears
        f += 4 * (S + L) ** 2 * D                             # joint attention
        if dual:
            f += 4 * S * S * D
        f += 2 * D * ((9 if dual else 6) + (2 if last else 6)) * D   # AdaLN linears
    vt_lin = 56 * S * D * D
    f += n_crossview * (vt_lin + 4 * seq_crossview * D * S)
    f += n_temporal * (vt_lin + 4 * seq_temporal * D * S)
    f += 2 * S * patch_k * D + 2 * S * D * out_n              # patch embed, proj_out
    f += 2 * D * 2 * D + 2 * (256 * D + D * D)                 # norm_out, timestep MLP
    if include_context_embed:
        f += 2 * L * joint_attention_dim * D
    return f


def gemm_flops(M, N, K):
    return 2.0 * M * N * K


if __name__ == "__main__":
    per_item = dit_flops_per_item()
    print("F_item = %.1f GFLOP" % (per_item / 1e9))
    print("F_step(N=192) = %.1f TFLOP" % (per_item * 192 / 1e12))
    print("excl. context_embedder: %.1f TFLOP" % (
        dit_flops_per_item(include_context_embed=False) * 192 / 1e12))
