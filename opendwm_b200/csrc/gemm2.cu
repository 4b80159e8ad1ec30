// 2-CTA (cta_group::2) variant of the persistent tcgen05 GEMM: a cluster of two CTAs on
// one TPC computes a 256x256 output tile with tcgen05.mma M=256.  Each CTA stages its own
// 128 rows of A and HALF of the W tile (128 of the 256 N rows), so per MMA every SM reads
// 16 KB (A) + 16 KB (W half) from shared memory instead of 16 + 32 KB — the kernel is
// power-capped on B200, so the saved shared-memory / L2 traffic turns into clock.
//
// Protocol (barriers live at identical offsets in both CTAs):
//   full[s]    leader only; both CTAs' TMA loads complete_tx on the LEADER's barrier
//              (cp.async.bulk.tensor ... .cta_group::2), the leader's producer expects 64 KB
//   empty[s]   both CTAs; the leader's MMA thread commits with a cluster multicast
//   tfull[a]   both CTAs (multicast commit) -> each CTA drains its own 128 TMEM lanes
//   tempty[a]  leader only; all 16 epilogue warps of the pair arrive (remote arrive)
#include "gemm_epilogue.cuh"

namespace dwm {

constexpr int G2_STAGES = 6;
constexpr int G2_A_BYTES = 128 * BK * 2;
constexpr int G2_B_BYTES = 128 * BK * 2;
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_SMEM_BYTES = G2_STAGES * G2_STAGE_BYTES + EPI_STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                void* smem, int32_t c0, int32_t c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1),
        "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit to the same barrier offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

template <typename T, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
    gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         int M, int N, int K, EpiParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + G2_STAGES * G2_A_BYTES;
  float4* epi_stage = reinterpret_cast<float4*>(smem + G2_STAGES * G2_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE_BYTES + EPI_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + G2_STAGES;
  uint64_t* tfull_bar = bars + 2 * G2_STAGES;
  uint64_t* tempty_bar = bars + 2 * G2_STAGES + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * G2_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  const int m_blocks = (M + 2 * BM - 1) / (2 * BM);
  const int n_blocks = (N + BN - 1) / BN;
  const int k_blocks = (K + BK - 1) / BK;
  const int num_tiles = m_blocks * n_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();   // barriers of both CTAs initialised, TMEM of the pair allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        const int m_blk = tile / n_blocks;
        const int n_blk = tile % n_blocks;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);
          tma_load_2d_2sm(&tmap_a, leader_full, smem_a + stage * G2_A_BYTES, kb * BK,
                          m_blk * 2 * BM + static_cast<int>(rank) * BM, kEvictNormal);
          tma_load_2d_2sm(&tmap_b, leader_full, smem_b + stage * G2_B_BYTES, kb * BK,
                          n_blk * BN + static_cast<int>(rank) * (BN / 2), kEvictLast);
          if (++stage == G2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc(2 * BM, BN, Cvt<T>::kUmmaFmt);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(smem_u32(smem_a + stage * G2_A_BYTES));
          const uint64_t db = umma_desc_sw128(smem_u32(smem_b + stage * G2_B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_f16_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit_2sm_mc(&empty_bar[stage]);
          if (++stage == G2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm_mc(&tfull_bar[as]);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..9 of both CTAs) =====================
    const int quarter = warp & 3;
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++it) {
      const int m_blk = tile / n_blocks;
      const int n_blk = tile % n_blocks;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      prefetch_resid_tile<EPI>(p, m_blk * 2 * BM + static_cast<int>(rank) * BM + quarter * 32 + lane, M, n_blk * BN, N, (warp - 2) >> 2);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(quarter * 32) << 16);
      drain_tile<T, EPI>(taddr, epi_stage + (warp - 2) * 256,
                         m_blk * 2 * BM + static_cast<int>(rank) * BM, quarter * 32, M, n_blk * BN, N, p, lane,
                         (warp - 2) >> 2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&tempty_bar[as]), 0));
    }
  }

  tc_fence_before();
  cluster_sync_all();   // nobody leaves while the peer may still touch its smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

template <typename T, int EPI>
static int launch_gemm2(const dwm_linear_args* a, cudaStream_t stream) {
  CUtensorMap ta, tb;
  int rc = make_tmap_2d(&ta, a->A, a->M, a->K, a->lda, BM, BK, 2);
  if (rc) return rc;
  rc = make_tmap_2d(&tb, a->W, a->N, a->K, a->ldw, BN / 2, BK, 2);
  if (rc) return rc;
  EpiParams p;
  fill_epi_params(p, a);
  auto kern = gemm2_tcgen05_kernel<T, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    DWM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES));
    attr_set = true;
  }
  const long long m_blocks = (a->M + 2 * BM - 1) / (2 * BM);
  const long long n_blocks = (a->N + BN - 1) / BN;
  const long long tiles = m_blocks * n_blocks;
  const int pairs = sm_count() / 2;
  const int clusters = static_cast<int>(tiles < pairs ? tiles : pairs);
  kern<<<2 * clusters, GEMM_THREADS, G2_SMEM_BYTES, stream>>>(ta, tb, static_cast<int>(a->M), static_cast<int>(a->N),
                                                             static_cast<int>(a->K), p);
  DWM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
static int dispatch2(const dwm_linear_args* a, cudaStream_t s) {
  switch (a->epilogue) {
    case DWM_EPI_STORE: return launch_gemm2<T, DWM_EPI_STORE>(a, s);
    case DWM_EPI_GEGLU: return launch_gemm2<T, DWM_EPI_GEGLU>(a, s);
    case DWM_EPI_QKNORM: return launch_gemm2<T, DWM_EPI_QKNORM>(a, s);
    case DWM_EPI_RESID: return launch_gemm2<T, DWM_EPI_RESID>(a, s);
    case DWM_EPI_F32: return launch_gemm2<T, DWM_EPI_F32>(a, s);
    default: set_last_error("dwm_b200_linear: unknown epilogue %d", a->epilogue); return -1;
  }
}

int gemm2_launch(const dwm_linear_args* a, cudaStream_t s) {
  if (a->dtype == DWM_BF16) return dispatch2<__nv_bfloat16>(a, s);
  if (a->dtype == DWM_F16) return dispatch2<__half>(a, s);
  set_last_error("dwm_b200_linear: dtype must be DWM_BF16 or DWM_F16, got %d", a->dtype);
  return -1;
}

}  // namespace dwm
