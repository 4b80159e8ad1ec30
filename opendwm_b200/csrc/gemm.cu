// Persistent warp-specialised tcgen05 GEMM with fused epilogues (sm_100a).
//
//   D[M,N] = A[M,K] · W[N,K]^T     A, W 16-bit (bf16/fp16), fp32 accumulation in TMEM
//
// Roles (192 threads, one CTA per SM, persistent over output tiles):
//   warp 0    TMA producer: 128x64 A tile + 256x64 W tile per k-block into a
//             4-stage 128B-swizzled shared-memory ring (mbarrier tx-count completion)
//   warp 1    MMA issuer: one elected thread issues tcgen05.mma 128x256x16, commits
//             stage release and accumulator-ready to mbarriers; owns TMEM alloc
//   warps 2-5 epilogue: tcgen05.ld the 128x256 fp32 accumulator (lane = row), apply
//             the fused epilogue and store; double-buffered TMEM (2 x 256 columns)
//             so the epilogue of tile i overlaps the main loop of tile i+1
//
// Replaces the cuBLASLt GEMM + ~10 elementwise launches per sub-layer that the
// reference runs.
#include "common.cuh"
#include "../../include/dwm_b200.h"

namespace dwm {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int B_STAGE_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int GEMM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int GEMM_SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct EpiParams {
  void* out;
  long long ldo;
  const float* bias;
  int act;
  long long rows_per_item, out_item_stride, out_row_offset;
  const float* qw;
  const float* kw;
  long long qk_region;
  float eps;
  const float* resid;
  long long ldr, resid_row_mod;
  const float* gate;
  long long gate_ld;
  const float* blend_x;
  long long ldx;
  const float* alpha;
  long long rows_per_batch;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case DWM_ACT_GELU_TANH: return gelu_tanh(v);
    case DWM_ACT_GELU_ERF: return gelu_erf(v);
    case DWM_ACT_SILU: return silu(v);
    default: return v;
  }
}

__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

// ---- per-epilogue tile drains (called by the 4 epilogue warps) -----------------
// `taddr` already contains this warp's lane quarter and the accumulator stage column.
// `m` is this thread's global row, `n_tile0` the first global column of the tile.

template <typename T>
__device__ __forceinline__ void epi_store(uint32_t taddr, int m, int M, int n_tile0, int N,
                                          const EpiParams& p) {
  long long orow = m;
  if (p.rows_per_item > 0) {
    orow = (m / p.rows_per_item) * p.out_item_stride + (m % p.rows_per_item);
  }
  orow += p.out_row_offset;
  T* out = reinterpret_cast<T*>(p.out) + orow * p.ldo;
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    const int n0 = n_tile0 + c * 32;
    if (n0 >= N) break;
    uint32_t r[32];
    tmem_ld32(taddr + c * 32, r);
    tmem_ld_wait();
    if (m < M) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float a = __uint_as_float(r[2 * j]);
        float b = __uint_as_float(r[2 * j + 1]);
        if (p.bias) {
          a += __ldg(p.bias + n0 + 2 * j);
          b += __ldg(p.bias + n0 + 2 * j + 1);
        }
        a = apply_act(a, p.act);
        b = apply_act(b, p.act);
        pk[j] = Cvt<T>::pack2(a, b);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        st_global_v4(out + n0 + 8 * j, pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
  }
}

template <typename T>
__device__ __forceinline__ void epi_geglu(uint32_t taddr, int m, int M, int n_tile0, int N,
                                          const EpiParams& p) {
  // tile columns [0,128) hold the value half, [128,256) the gate half of output
  // columns [n_tile0/2, n_tile0/2 + 128).
  T* out = reinterpret_cast<T*>(p.out) + static_cast<long long>(m) * p.ldo + n_tile0 / 2;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t rv[32], rg[32];
    tmem_ld32(taddr + c * 32, rv);
    tmem_ld32(taddr + 128 + c * 32, rg);
    tmem_ld_wait();
    if (m < M) {
      const float* bv = p.bias ? p.bias + n_tile0 + c * 32 : nullptr;
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float v0 = __uint_as_float(rv[2 * j]), v1 = __uint_as_float(rv[2 * j + 1]);
        float g0 = __uint_as_float(rg[2 * j]), g1 = __uint_as_float(rg[2 * j + 1]);
        if (bv) {
          v0 += __ldg(bv + 2 * j);
          v1 += __ldg(bv + 2 * j + 1);
          g0 += __ldg(bv + 128 + 2 * j);
          g1 += __ldg(bv + 128 + 2 * j + 1);
        }
        pk[j] = Cvt<T>::pack2(v0 * gelu_erf(g0), v1 * gelu_erf(g1));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        st_global_v4(out + c * 32 + 8 * j, pk[4 * j], pk[4 * j + 1], pk[4 * j + 2],
                     pk[4 * j + 3]);
    }
  }
}

template <typename T>
__device__ __forceinline__ void epi_qknorm(uint32_t taddr, int m, int M, int n_tile0, int N,
                                           const EpiParams& p) {
  long long orow = m;
  if (p.rows_per_item > 0) {
    orow = (m / p.rows_per_item) * p.out_item_stride + (m % p.rows_per_item);
  }
  orow += p.out_row_offset;
  T* out = reinterpret_cast<T*>(p.out) + orow * p.ldo;
#pragma unroll 1
  for (int g = 0; g < BN / 64; ++g) {
    const int n0 = n_tile0 + g * 64;
    if (n0 >= N) break;
    uint32_t r0[32], r1[32];
    tmem_ld32(taddr + g * 64, r0);
    tmem_ld32(taddr + g * 64 + 32, r1);
    tmem_ld_wait();
    if (m < M) {
      float v[64];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = __uint_as_float(r0[j]);
        v[32 + j] = __uint_as_float(r1[j]);
      }
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] += __ldg(p.bias + n0 + j);
      }
      const int region = n0 / static_cast<int>(p.qk_region);
      if (region < 2) {
        const float* w = region == 0 ? p.qw : p.kw;
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) ss += v[j] * v[j];
        const float inv = rsqrtf(ss * (1.0f / 64.0f) + p.eps);
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = v[j] * inv * __ldg(w + j);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        st_global_v4(out + n0 + 8 * j, Cvt<T>::pack2(v[8 * j], v[8 * j + 1]),
                     Cvt<T>::pack2(v[8 * j + 2], v[8 * j + 3]),
                     Cvt<T>::pack2(v[8 * j + 4], v[8 * j + 5]),
                     Cvt<T>::pack2(v[8 * j + 6], v[8 * j + 7]));
    }
  }
}

template <bool kResid>
__device__ __forceinline__ void epi_f32(uint32_t taddr, int m, int M, int n_tile0, int N,
                                        const EpiParams& p) {
  float* out = reinterpret_cast<float*>(p.out) + static_cast<long long>(m) * p.ldo;
  const float* gate = nullptr;
  const float* resid = nullptr;
  const float* bx = nullptr;
  float alpha = 0.f;
  if (kResid && m < M) {
    if (p.gate) {
      long long item = p.rows_per_item > 0 ? m / p.rows_per_item : 0;
      gate = p.gate + item * p.gate_ld;
    }
    if (p.resid) {
      long long rr = p.resid_row_mod > 0 ? m % p.resid_row_mod : m;
      resid = p.resid + rr * p.ldr;
    }
    if (p.blend_x) {
      bx = p.blend_x + static_cast<long long>(m) * p.ldx;
      alpha = __ldg(p.alpha + (p.rows_per_batch > 0 ? m / p.rows_per_batch : 0));
    }
  }
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    const int n0 = n_tile0 + c * 32;
    if (n0 >= N) break;
    uint32_t r[32];
    tmem_ld32(taddr + c * 32, r);
    tmem_ld_wait();
    if (m < M) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 v;
        v.x = __uint_as_float(r[4 * j]);
        v.y = __uint_as_float(r[4 * j + 1]);
        v.z = __uint_as_float(r[4 * j + 2]);
        v.w = __uint_as_float(r[4 * j + 3]);
        if (p.bias) {
          float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + 4 * j));
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (kResid) {
          if (gate) {
            float4 g = __ldg(reinterpret_cast<const float4*>(gate + n0 + 4 * j));
            v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
          }
          if (resid) {
            float4 q = *reinterpret_cast<const float4*>(resid + n0 + 4 * j);
            v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
          }
          if (bx) {
            float4 x = *reinterpret_cast<const float4*>(bx + n0 + 4 * j);
            const float b1 = 1.0f - alpha;
            v.x = alpha * x.x + b1 * v.x;
            v.y = alpha * x.y + b1 * v.y;
            v.z = alpha * x.z + b1 * v.z;
            v.w = alpha * x.w + b1 * v.w;
          }
        } else {
          v.x = apply_act(v.x, p.act);
          v.y = apply_act(v.y, p.act);
          v.z = apply_act(v.z, p.act);
          v.w = apply_act(v.w, p.act);
        }
        *reinterpret_cast<float4*>(out + n0 + 4 * j) = v;
      }
    }
  }
}

template <typename T, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                        const __grid_constant__ CUtensorMap tmap_b, int M, int N, int K,
                        EpiParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled UMMA/TMA tiles need 1024-byte alignment.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]       MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]   epilogue -> MMA
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_blocks = (M + BM - 1) / BM;
  const int n_blocks = (N + BN - 1) / BN;
  const int k_blocks = (K + BK - 1) / BK;
  const int num_tiles = m_blocks * n_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_blocks;
        const int n_blk = tile % n_blocks;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(&tmap_a, &full_bar[stage], smem_a + stage * A_STAGE_BYTES, kb * BK,
                      m_blk * BM, kEvictNormal);
          tma_load_2d(&tmap_b, &full_bar[stage], smem_b + stage * B_STAGE_BYTES, kb * BK,
                      n_blk * BN, kEvictLast);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(BM, BN, Cvt<T>::kUmmaFmt);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(smem_u32(smem_a + stage * A_STAGE_BYTES));
          const uint64_t db = umma_desc_sw128(smem_u32(smem_b + stage * B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the 128B swizzle atom
            umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / n_blocks;
      const int n_blk = tile % n_blocks;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(quarter * 32) << 16);
      const int m = m_blk * BM + quarter * 32 + lane;
      const int n_tile0 = n_blk * BN;
      if constexpr (EPI == DWM_EPI_STORE) epi_store<T>(taddr, m, M, n_tile0, N, p);
      if constexpr (EPI == DWM_EPI_GEGLU) epi_geglu<T>(taddr, m, M, n_tile0, N, p);
      if constexpr (EPI == DWM_EPI_QKNORM) epi_qknorm<T>(taddr, m, M, n_tile0, N, p);
      if constexpr (EPI == DWM_EPI_RESID) epi_f32<true>(taddr, m, M, n_tile0, N, p);
      if constexpr (EPI == DWM_EPI_F32) epi_f32<false>(taddr, m, M, n_tile0, N, p);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <typename T, int EPI>
static int launch_gemm(const dwm_linear_args* a, cudaStream_t stream) {
  CUtensorMap ta, tb;
  int rc = make_tmap_2d(&ta, a->A, a->M, a->K, a->lda, BM, BK, 2);
  if (rc) return rc;
  rc = make_tmap_2d(&tb, a->W, a->N, a->K, a->ldw, BN, BK, 2);
  if (rc) return rc;

  EpiParams p;
  p.out = a->out;
  p.ldo = a->ldo;
  p.bias = a->bias;
  p.act = a->act;
  p.rows_per_item = a->rows_per_item;
  p.out_item_stride = a->out_item_stride;
  p.out_row_offset = a->out_row_offset;
  p.qw = a->q_norm_weight;
  p.kw = a->k_norm_weight;
  p.qk_region = a->qk_region;
  p.eps = a->eps;
  p.resid = a->resid;
  p.ldr = a->ldr;
  p.resid_row_mod = a->resid_row_mod;
  p.gate = a->gate;
  p.gate_ld = a->gate_ld;
  p.blend_x = a->blend_x;
  p.ldx = a->ldx;
  p.alpha = a->alpha;
  p.rows_per_batch = a->rows_per_batch;

  auto kern = gemm_tcgen05_kernel<T, EPI>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    DWM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        GEMM_SMEM_BYTES));
    attr_set = true;
  }
  const long long m_blocks = (a->M + BM - 1) / BM;
  const long long n_blocks = (a->N + BN - 1) / BN;
  const long long tiles = m_blocks * n_blocks;
  const int sms = sm_count();
  const int grid = static_cast<int>(tiles < sms ? tiles : sms);
  kern<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(ta, tb, static_cast<int>(a->M),
                                                        static_cast<int>(a->N),
                                                        static_cast<int>(a->K), p);
  DWM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
static int dispatch_epi(const dwm_linear_args* a, cudaStream_t s) {
  switch (a->epilogue) {
    case DWM_EPI_STORE: return launch_gemm<T, DWM_EPI_STORE>(a, s);
    case DWM_EPI_GEGLU: return launch_gemm<T, DWM_EPI_GEGLU>(a, s);
    case DWM_EPI_QKNORM: return launch_gemm<T, DWM_EPI_QKNORM>(a, s);
    case DWM_EPI_RESID: return launch_gemm<T, DWM_EPI_RESID>(a, s);
    case DWM_EPI_F32: return launch_gemm<T, DWM_EPI_F32>(a, s);
    default: set_last_error("dwm_b200_linear: unknown epilogue %d", a->epilogue); return -1;
  }
}

}  // namespace dwm

extern "C" int dwm_b200_linear(const dwm_linear_args* a, dwm_stream_t stream) {
  using namespace dwm;
  DWM_REQUIRE(a != nullptr, "dwm_b200_linear: null args");
  DWM_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "dwm_b200_linear: empty problem %lld x %lld x %lld",
              (long long)a->M, (long long)a->N, (long long)a->K);
  DWM_REQUIRE(a->M < (1ll << 31) && a->N < (1ll << 31) && a->K < (1ll << 31),
              "dwm_b200_linear: dimension exceeds int32");
  DWM_REQUIRE(a->A && a->W && a->out, "dwm_b200_linear: null A/W/out");
  DWM_REQUIRE(a->K % 8 == 0 && a->lda % 8 == 0 && a->ldw % 8 == 0,
              "dwm_b200_linear: K, lda, ldw must be multiples of 8 (16-byte TMA pitch); got %lld %lld %lld",
              (long long)a->K, (long long)a->lda, (long long)a->ldw);
  DWM_REQUIRE((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(a->W) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
              "dwm_b200_linear: A, W, out must be 16-byte aligned");
  DWM_REQUIRE(a->N % 32 == 0, "dwm_b200_linear: N must be a multiple of 32, got %lld", (long long)a->N);
  DWM_REQUIRE(a->ldo % 8 == 0, "dwm_b200_linear: ldo must be a multiple of 8");
  if (a->epilogue == DWM_EPI_GEGLU)
    DWM_REQUIRE(a->N % 256 == 0, "dwm_b200_linear: GEGLU needs N %% 256 == 0 (packed weight)");
  if (a->epilogue == DWM_EPI_QKNORM) {
    DWM_REQUIRE(a->N % 64 == 0 && a->qk_region > 0 && a->qk_region % 64 == 0 && a->q_norm_weight &&
                    a->k_norm_weight,
                "dwm_b200_linear: QKNORM needs head_dim 64 regions and both norm weights");
  }
  if (a->epilogue == DWM_EPI_RESID && a->blend_x)
    DWM_REQUIRE(a->alpha != nullptr, "dwm_b200_linear: blend_x without alpha");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (a->dtype == DWM_BF16) return dispatch_epi<__nv_bfloat16>(a, s);
  if (a->dtype == DWM_F16) return dispatch_epi<__half>(a, s);
  set_last_error("dwm_b200_linear: dtype must be DWM_BF16 or DWM_F16, got %d", a->dtype);
  return -1;
}
