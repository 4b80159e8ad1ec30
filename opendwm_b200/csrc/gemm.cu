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
#include <stdlib.h>
#include <string.h>

#include "gemm_epilogue.cuh"

namespace dwm {

constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int B_STAGE_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int GEMM_SMEM_BYTES =
    STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

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
      prefetch_resid_tile<EPI>(p, m_blk * BM + quarter * 32 + lane, M, n_blk * BN, N, (warp - 2) >> 2);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(quarter * 32) << 16);
      drain_tile<T, EPI>(taddr, epi_stage + (warp - 2) * 256, m_blk * BM, quarter * 32, M,
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
  fill_epi_params(p, a);

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

int gemm2_launch(const dwm_linear_args* a, cudaStream_t s);   // gemm2.cu
int g_gemm_2cta = -1;   // -1: from env DWM_GEMM_2CTA (default 1), 0 / 1: forced

}  // namespace dwm

namespace dwm { int g_attn_tc = -1; extern int g_ln_staged; extern int g_resid_tma; extern int g_gemm_bn; extern int g_conv_2cta; extern int g_conv_halo; }

extern "C" int dwm_b200_set_option(const char* name, int value) {
  using namespace dwm;
  DWM_REQUIRE(name != nullptr, "dwm_b200_set_option: null name");
  if (strcmp(name, "gemm_2cta") == 0) { g_gemm_2cta = value; return 0; }
  if (strcmp(name, "attn_tc") == 0) { g_attn_tc = value; return 0; }
  if (strcmp(name, "ln_staged") == 0) { g_ln_staged = value; return 0; }
  if (strcmp(name, "resid_tma") == 0) { g_resid_tma = value; return 0; }
  if (strcmp(name, "gemm_bn") == 0) { g_gemm_bn = value; return 0; }
  if (strcmp(name, "conv_2cta") == 0) { g_conv_2cta = value; return 0; }
  if (strcmp(name, "conv_halo") == 0) { g_conv_halo = value; return 0; }
  set_last_error("dwm_b200_set_option: unknown option %s", name);
  return -1;
}

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
  DWM_REQUIRE(a->n_peer_out >= 0 && a->n_peer_out <= 8, "dwm_b200_linear: n_peer_out out of range");
  if (a->n_peer_out > 0)
    DWM_REQUIRE(a->epilogue == DWM_EPI_STORE || a->epilogue == DWM_EPI_QKNORM || a->epilogue == DWM_EPI_GEGLU,
                "dwm_b200_linear: peer_out needs a 16-bit epilogue");
  if (a->epilogue == DWM_EPI_RESID && a->blend_x)
    DWM_REQUIRE(a->alpha != nullptr, "dwm_b200_linear: blend_x without alpha");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (g_gemm_2cta < 0) {
    const char* e = getenv("DWM_GEMM_2CTA");
    g_gemm_2cta = (e && e[0] == '0') ? 0 : 1;
  }
  // 2-CTA clusters pay off once there are enough 256-row tiles to fill the 74 SM pairs
  if (g_gemm_2cta == 1 && a->M >= 512) return gemm2_launch(a, s);
  if (a->dtype == DWM_BF16) return dispatch_epi<__nv_bfloat16>(a, s);
  if (a->dtype == DWM_F16) return dispatch_epi<__half>(a, s);
  set_last_error("dwm_b200_linear: dtype must be DWM_BF16 or DWM_F16, got %d", a->dtype);
  return -1;
}
