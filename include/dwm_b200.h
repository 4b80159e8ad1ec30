/*
 * dwm_b200 — C ABI of the B200-native (sm_100a) kernels behind OpenDWM's CTSD
 * denoising hot path.
 *
 * The reference (SenseTime-FVG/OpenDWM) has no C/FFI interface of its own: its
 * plug point is the JSON `_class_name` factory (src/dwm/common.py:133-172) that
 * instantiates `dwm.models.crossview_temporal_dit.DiTCrossviewTemporalConditionModel`
 * etc.  The Python mirror of those classes (src/dwm in this repo) calls the
 * entry points below through ctypes.  Each entry point cites the reference
 * code whose arithmetic it replaces.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error; the message is available
 *    from dwm_b200_last_error() (thread local).
 *  - all pointers are DEVICE pointers owned by the caller (PyTorch); nothing is
 *    allocated here.  Launches are asynchronous on `stream`.
 *  - 16-bit activations/weights are bf16 (DWM_BF16) or fp16 (DWM_F16);
 *    biases, norm weights, modulation vectors, residual streams are fp32.
 *  - there is NO CPU fallback: calling these without a Blackwell GPU fails.
 */
#ifndef DWM_B200_H_
#define DWM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dwm_stream_t; /* cudaStream_t */

enum dwm_dtype { DWM_BF16 = 0, DWM_F16 = 1, DWM_F32 = 2 };
enum dwm_act { DWM_ACT_NONE = 0, DWM_ACT_GELU_TANH = 1, DWM_ACT_GELU_ERF = 2, DWM_ACT_SILU = 3 };

/* Epilogues of dwm_b200_linear (all fused into the tcgen05 GEMM kernel). */
enum dwm_epilogue {
  /* out16[r(m), n] = act(acc + bias[n]) */
  DWM_EPI_STORE = 0,
  /* GEGLU (diffusers FeedForward activation_fn="geglu", used by
   * VTSelfAttentionBlock.ff_in / .ff, crossview_temporal.py:548,559):
   * weight rows are pre-packed in blocks of 256 = [128 value rows | 128 gate rows];
   * out16[m, j] = (acc_v + b_v) * gelu_erf(acc_g + b_g), out width N/2. */
  DWM_EPI_GEGLU = 1,
  /* fused q/k/v projection + per-head RMSNorm(q), RMSNorm(k) (head_dim 64;
   * diffusers Attention qk_norm="rms_norm", crossview_temporal.py:552-555):
   * columns [0,D) are q, [D,2D) k, [2D,3D) v with D = qk_region. */
  DWM_EPI_QKNORM = 2,
  /* fp32 residual stream update, optionally gated and alpha-blended:
   *   v = acc + bias[n];  v *= gate[item(m), n];  v += resid[rr(m), n];
   *   if blend_x: v = alpha[b(m)] * blend_x[m, n] + (1 - alpha[b(m)]) * v
   *   out32[m, n] = v
   * (JointTransformerBlock gated residuals; AlphaBlender crossview_temporal.py:53-72) */
  DWM_EPI_RESID = 3,
  /* out32[m, n] = act(acc + bias[n]) */
  DWM_EPI_F32 = 4
};

typedef struct dwm_linear_args {
  int64_t M, N, K;
  const void* A;   /* [M, K] 16-bit, row pitch lda elements */
  int64_t lda;
  const void* W;   /* [N, K] 16-bit (torch.nn.Linear weight layout), row pitch ldw */
  int64_t ldw;
  const float* bias; /* [N] or NULL */
  int dtype;       /* dwm_dtype of A, W and of 16-bit outputs */
  int epilogue;    /* dwm_epilogue */
  int act;         /* dwm_act (STORE / F32 epilogues) */
  void* out;
  int64_t ldo;     /* row pitch of out, elements */
  /* item structure of the M rows: item(m) = m / rows_per_item (0 => one item).
   * 16-bit outputs go to row  item * out_item_stride + m % rows_per_item + out_row_offset
   * (lets sample and context tokens of one view-frame land in one joint buffer). */
  int64_t rows_per_item;
  int64_t out_item_stride;
  int64_t out_row_offset;
  /* QKNORM */
  const float* q_norm_weight; /* [64] */
  const float* k_norm_weight; /* [64] */
  int64_t qk_region;          /* D */
  float eps;
  /* RESID */
  const float* resid;   /* fp32 [*, ldr] or NULL */
  int64_t ldr;
  int64_t resid_row_mod; /* rr(m) = resid_row_mod ? m % resid_row_mod : m */
  const float* gate;    /* fp32 [items, gate_ld] or NULL */
  int64_t gate_ld;
  const float* blend_x; /* fp32 [M, ldx] or NULL */
  int64_t ldx;
  const float* alpha;   /* fp32 [batches]; b(m) = m / rows_per_batch */
  int64_t rows_per_batch;
} dwm_linear_args;

const char* dwm_b200_version(void);
const char* dwm_b200_last_error(void);

/* y = epilogue(A @ W^T): replaces every torch.nn.Linear / 1x1 / patchify conv on the
 * path (diffusers Attention.to_q/k/v/to_out, FeedForward, AdaLayerNormZero.linear,
 * PatchEmbed.proj, SD3Transformer2DModel.proj_out; call sites
 * crossview_temporal_dit.py:421-431,517-521,599-600, crossview_temporal.py:562-582). */
int dwm_b200_linear(const dwm_linear_args* args, dwm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DWM_B200_H_ */
