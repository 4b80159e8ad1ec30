// Persistent warp-specialised tcgen05 GEMM with fused epilogues (sm_100a).
//
//   D[M,N] = A[M,K] · W[N,K]^T     A, W 16-bit (bf16/fp16), fp32 accumulation in TMEM
//
// Roles (320 threads, one CTA per SM, persistent over output tiles):
//   warp 0    TMA producer: 128x64 A tile + 256x64 W tile per k-block into a
//             4-stage 128B-swizzled shared-memory ring (mbarrier tx-count completion)
//   warp 1    MMA issuer: one elected thread issues tcgen05.mma 128x256x16, commits
//             stage release and accumulator-ready to mbarriers; owns TMEM alloc
//   warps 2-9 epilogue: tcgen05.ld the 128x256 fp32 accumulator (lane = row; two warps
//             per TMEM lane quarter, interleaved over 32-column chunks), apply the fused
//             epilogue and store; double-buffered TMEM (2 x 256 columns) so the
//             epilogue of tile i overlaps the main loop of tile i+1
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
constexpr int EPI_WARPS = 8;  // two warps per TMEM lane quarter, interleaved over column chunks
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int EPI_STAGE_BYTES = EPI_WARPS * 32 * 32 * 4;  // per epilogue warp: 32x32 fp32
constexpr int GEMM_SMEM_BYTES =
    STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

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
  int norm_regions;
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
    case DWM_ACT_RELU: return fmaxf(v, 0.f);
    default: return v;
  }
}

__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

// ---- tile drain (4 epilogue warps) ----------------------------------------------
// Each warp owns 32 accumulator rows (TMEM lane quarter).  tcgen05.ld hands every
// thread one ROW, which would make global accesses 32-way scattered.  So each
// 32x32 fp32 chunk is transposed through a 4 KB XOR-swizzled shared-memory
// staging buffer: phase 1 (thread = row) applies the row-local math and dumps,
// phase 2 (8 lanes per row, float4 per lane, 4 rows per instruction) does the
// coalesced global traffic (incl. the fp32 residual read-modify-write).

__device__ __forceinline__ void stage_dump(float4* stg, int lane, const float (&v)[32]) {
  __syncwarp();  // phase-2 readers of the previous chunk are done
#pragma unroll
  for (int j = 0; j < 8; ++j)
    stg[lane * 8 + (j ^ (lane & 7))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
}

template <typename T, int EPI>
__device__ __forceinline__ void drain_tile(uint32_t taddr, float4* stg, int m0, int M,
                                           int n_tile0, int N, const EpiParams& p, int lane,
                                           int half) {
  constexpr bool kOut16 = (EPI == DWM_EPI_STORE || EPI == DWM_EPI_GEGLU || EPI == DWM_EPI_QKNORM);
  const int rs = lane >> 3;  // phase-2: row within a group of 4
  const int c4 = lane & 7;   // phase-2: float4 column within the 32-col chunk

  // phase-2 per-row metadata for the 8 rows this lane stores (it*4 + rs)
  int orow[8];
  int rrow[8];
  int item[8];
  float alpha[8];
  const int rpi = static_cast<int>(p.rows_per_item);
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int m = m0 + it * 4 + rs;
    if (m < M) {
      int o = m;
      if (kOut16) {
        if (rpi > 0) o = (m / rpi) * static_cast<int>(p.out_item_stride) + (m % rpi);
        o += static_cast<int>(p.out_row_offset);
      }
      orow[it] = o;
      if (EPI == DWM_EPI_RESID) {
        item[it] = rpi > 0 ? m / rpi : 0;
        rrow[it] = p.resid_row_mod > 0 ? m % static_cast<int>(p.resid_row_mod) : m;
        alpha[it] = p.blend_x ? __ldg(p.alpha + (p.rows_per_batch > 0 ? m / static_cast<int>(p.rows_per_batch) : 0)) : 0.f;
      }
    } else {
      orow[it] = -1;
    }
  }

  // phase 2 for 16-bit outputs: `ocol` = first output column of the staged chunk
  auto flush16 = [&](int ocol) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = it * 4 + rs;
      const float4 v = stg[r * 8 + (c4 ^ (r & 7))];
      if (orow[it] >= 0) {
        T* dst = reinterpret_cast<T*>(p.out) + static_cast<long long>(orow[it]) * p.ldo + ocol + c4 * 4;
        uint2 pk;
        pk.x = Cvt<T>::pack2(v.x, v.y);
        pk.y = Cvt<T>::pack2(v.z, v.w);
        *reinterpret_cast<uint2*>(dst) = pk;
      }
    }
  };
  // phase 2 for fp32 outputs (optionally gated / residual / blended).  The residual
  // (and blend) operands are prefetched into registers at the top of each chunk so
  // their HBM latency overlaps the TMEM load + transpose; in-place update is safe
  // because each lane reads exactly the elements it later writes.
  float4 rq[8], bq[8];
  auto prefetch32 = [&](int ocol) {
    if constexpr (EPI == DWM_EPI_RESID) {
      const int col = ocol + c4 * 4;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        rq[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        bq[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (orow[it] >= 0) {
          if (p.resid) rq[it] = *reinterpret_cast<const float4*>(p.resid + static_cast<long long>(rrow[it]) * p.ldr + col);
          if (p.blend_x) bq[it] = *reinterpret_cast<const float4*>(p.blend_x + static_cast<long long>(orow[it]) * p.ldx + col);
        }
      }
    }
  };
  auto flush32 = [&](int ocol) {
    const int col = ocol + c4 * 4;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (EPI == DWM_EPI_RESID && p.bias) b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = it * 4 + rs;
      float4 v = stg[r * 8 + (c4 ^ (r & 7))];
      if (orow[it] >= 0) {
        if (EPI == DWM_EPI_RESID) {
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
          if (p.gate) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.gate + static_cast<long long>(item[it]) * p.gate_ld + col));
            v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
          }
          v.x += rq[it].x; v.y += rq[it].y; v.z += rq[it].z; v.w += rq[it].w;
          if (p.blend_x) {
            const float a = alpha[it], a1 = 1.0f - alpha[it];
            v.x = a * bq[it].x + a1 * v.x; v.y = a * bq[it].y + a1 * v.y;
            v.z = a * bq[it].z + a1 * v.z; v.w = a * bq[it].w + a1 * v.w;
          }
        }
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<long long>(orow[it]) * p.ldo + col) = v;
      }
    }
  };

  if constexpr (EPI == DWM_EPI_STORE || EPI == DWM_EPI_F32 || EPI == DWM_EPI_RESID) {
#pragma unroll 1
    for (int c = half; c < BN / 32; c += 2) {
      const int n0 = n_tile0 + c * 32;
      if (n0 >= N) break;
      prefetch32(n0);
      uint32_t r[32];
      tmem_ld32(taddr + c * 32, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (EPI != DWM_EPI_RESID) {
        if (p.bias) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
          }
        }
        // activation selected once per chunk (warp-uniform), loops fully unrolled
        if (p.act == DWM_ACT_GELU_TANH) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(v[j]);
        } else if (p.act == DWM_ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = silu(v[j]);
        } else if (p.act == DWM_ACT_GELU_ERF) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        } else if (p.act == DWM_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
      }
      stage_dump(stg, lane, v);
      if constexpr (EPI == DWM_EPI_STORE) flush16(n0); else flush32(n0);
    }
  } else if constexpr (EPI == DWM_EPI_GEGLU) {
    // tile columns [0,128) hold the value half, [128,256) the gate half of output
    // columns [n_tile0/2, n_tile0/2 + 128).
#pragma unroll 1
    for (int c = half; c < 4; c += 2) {
      uint32_t rv[32], rg[32];
      tmem_ld32(taddr + c * 32, rv);
      tmem_ld32(taddr + 128 + c * 32, rg);
      tmem_ld_wait();
      const float* bv = p.bias ? p.bias + n_tile0 + c * 32 : nullptr;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float a = __uint_as_float(rv[j]);
        float g = __uint_as_float(rg[j]);
        if (bv) {
          a += __ldg(bv + j);
          g += __ldg(bv + 128 + j);
        }
        v[j] = a * gelu_erf(g);
      }
      stage_dump(stg, lane, v);
      flush16(n_tile0 / 2 + c * 32);
    }
  } else {  // DWM_EPI_QKNORM: 64-column heads
#pragma unroll 1
    for (int g = half; g < BN / 64; g += 2) {
      const int n0 = n_tile0 + g * 64;
      if (n0 >= N) break;
      uint32_t r0[32], r1[32];
      tmem_ld32(taddr + g * 64, r0);
      tmem_ld32(taddr + g * 64 + 32, r1);
      tmem_ld_wait();
      float v0[32], v1[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v0[j] = __uint_as_float(r0[j]);
        v1[j] = __uint_as_float(r1[j]);
      }
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v0[j] += __ldg(p.bias + n0 + j);
          v1[j] += __ldg(p.bias + n0 + 32 + j);
        }
      }
      const int region = n0 / static_cast<int>(p.qk_region);
      if (region < p.norm_regions) {
        const float* w = region == 0 ? p.qw : p.kw;
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) ss += v0[j] * v0[j] + v1[j] * v1[j];
        const float inv = rsqrtf(ss * (1.0f / 64.0f) + p.eps);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v0[j] = v0[j] * inv * __ldg(w + j);
          v1[j] = v1[j] * inv * __ldg(w + 32 + j);
        }
      }
      stage_dump(stg, lane, v0);
      flush16(n0);
      stage_dump(stg, lane, v1);
      flush16(n0 + 32);
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
  float4* epi_stage = reinterpret_cast<float4*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_STAGE_BYTES);
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
      mbar_init(&tempty_bar[s], EPI_WARPS);
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
    // ===================== epilogue (warps 2..9) =====================
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
      drain_tile<T, EPI>(taddr, epi_stage + (warp - 2) * 256, m_blk * BM + quarter * 32, M,
                         n_blk * BN, N, p, lane, (warp - 2) >> 2);
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
  p.norm_regions = a->qk_norm_regions > 0 ? a->qk_norm_regions : 2;
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
                    (a->k_norm_weight || a->qk_norm_regions == 1),
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
