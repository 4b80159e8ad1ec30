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
#include "cta_pair.cuh"

namespace dwm {

constexpr int G2_A_BYTES = 128 * BK * 2;
// RT ("residual through TMA") instantiations of the RESID epilogue stage the fp32 residual /
// result tile through shared memory with TMA in both directions (2 x 4 KB chunk buffers per
// epilogue warp = 64 KB) and run one operand stage less.
constexpr int RT_CHUNK_BYTES = 32 * 32 * 4;
constexpr int RT_EPI_BYTES = EPI_WARPS * 2 * RT_CHUNK_BYTES;
// BNT = accumulator columns of a tile: 256, or 128 for RT launches whose 256-wide tiling would
// leave a large tail wave on the 74 clusters (N = 1536 at M = 10 752 rows per rank: 252 tiles =
// 3.4 waves -> 504 tiles = 6.8 waves).  Per MMA step a CTA then reads 4 KB (A) + 2 KB (W half)
// per 64 cycles = 96 B/clk of shared memory instead of 64 B/clk.
template <bool RT, int BNT> struct G2Cfg {
  static constexpr int kBBytes = (BNT / 2) * BK * 2;
  static constexpr int kStageBytes = G2_A_BYTES + kBBytes;
  static constexpr int kStages = RT ? (BNT == 128 ? 6 : 5) : 6;
  static constexpr int kEpiBytes = RT ? RT_EPI_BYTES : EPI_STAGE_BYTES;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 + 512;
};

// ---- bulk-tensor store (shared -> global) and its group bookkeeping ------------------------
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// RESID epilogue, TMA in / TMA out (one epilogue warp, one 32-row x 32-column fp32 chunk):
//   thread = accumulator row (TMEM lane); the residual chunk was brought in by TMA with the
//   128-byte swizzle (16-byte unit j of row r sits at unit j ^ (r & 7)), so the 32 lanes read
//   and write their own rows conflict-free and NO transpose is needed; the result leaves through
//   a bulk tensor store, which also clips the M tail.
//     v = (acc + bias[n]) * gate[item(m), n] + resid[m, n]
//     blend: v = alpha[b(m)] * blend_x[m, n] + (1 - alpha[b(m)]) * v
// bias / gate values of one chunk, fetched into registers BEFORE the waits on the accumulator
// and on the residual chunk so that their L2 latency is hidden behind those waits (ncu r02: the
// gate multiplies were the top stall of the epilogue when loaded at the point of use)
struct RtVecs {
  float4 b[8];
  float4 g[8];
};
__device__ __forceinline__ void rt_load_vecs(RtVecs& v, const EpiParams& p, int n0, int item, bool live) {
  const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
  const float4* g4 = reinterpret_cast<const float4*>(p.gate + static_cast<long long>(item) * p.gate_ld + n0);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v.b[j] = (p.bias && live) ? __ldg(b4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    v.g[j] = (p.gate && live) ? __ldg(g4 + j) : make_float4(1.f, 1.f, 1.f, 1.f);
  }
}
__device__ __forceinline__ void rt_chunk_math(const uint32_t (&acc)[32], const RtVecs& vv, float* rbuf,
                                              const float* xbuf, const EpiParams& p, float alpha, int lane) {
  float4* row = reinterpret_cast<float4*>(rbuf) + lane * 8;
  const float4* xrow = xbuf ? reinterpret_cast<const float4*>(xbuf) + lane * 8 : nullptr;
  const float a1 = 1.0f - alpha;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int u = j ^ (lane & 7);
    float4 v = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                           __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
    const float4 b = vv.b[j], g = vv.g[j];      // 0 / 1 when absent (rt_load_vecs)
    const float4 r = row[u];
    v.x = resid_elem(v.x, b.x, g.x, r.x); v.y = resid_elem(v.y, b.y, g.y, r.y);
    v.z = resid_elem(v.z, b.z, g.z, r.z); v.w = resid_elem(v.w, b.w, g.w, r.w);
    if (xrow) {
      const float4 x = xrow[u];
      v.x = blend_elem(alpha, a1, x.x, v.x); v.y = blend_elem(alpha, a1, x.y, v.y);
      v.z = blend_elem(alpha, a1, x.z, v.z); v.w = blend_elem(alpha, a1, x.w, v.w);
    }
    row[u] = v;
  }
}

template <typename T, int EPI, bool RT, int BNT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
    gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_r, const __grid_constant__ CUtensorMap tmap_o,
                         const __grid_constant__ CUtensorMap tmap_x, int M, int N, int K, EpiParams p) {
  static_assert(!RT || EPI == DWM_EPI_RESID, "RT is a RESID epilogue variant");
  static_assert(BNT == 256 || (RT && BNT == 128), "narrow tiles exist for the RT epilogue only");
  using Cfg = G2Cfg<RT, BNT>;
  constexpr int G2_STAGES = Cfg::kStages;
  constexpr int G2_B_BYTES = Cfg::kBBytes;
  constexpr int G2_STAGE_BYTES = Cfg::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + G2_STAGES * G2_A_BYTES;
  float4* epi_stage = reinterpret_cast<float4*>(smem + G2_STAGES * G2_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE_BYTES + Cfg::kEpiBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + G2_STAGES;
  uint64_t* tfull_bar = bars + 2 * G2_STAGES;
  uint64_t* tempty_bar = bars + 2 * G2_STAGES + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * G2_STAGES + 4);
  uint64_t* rt_bar = bars + 2 * G2_STAGES + 6;   // RT: 2 per epilogue warp (chunk landed)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  const int m_blocks = (M + 2 * BM - 1) / (2 * BM);
  const int n_blocks = (N + BNT - 1) / BNT;
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
    if constexpr (RT) {
      tma_prefetch_desc(&tmap_r);
      tma_prefetch_desc(&tmap_o);
      if (p.blend_x) tma_prefetch_desc(&tmap_x);
      for (int s = 0; s < 2 * EPI_WARPS; ++s) mbar_init(&rt_bar[s], 1);
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
                          n_blk * BNT + static_cast<int>(rank) * (BNT / 2), kEvictLast);
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
      constexpr uint32_t idesc = umma_idesc(2 * BM, BNT, Cvt<T>::kUmmaFmt);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BNT;
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
  } else if constexpr (RT) {
    // ========== epilogue, residual tile through TMA (warps 2..9 of both CTAs) ==========
    // Each warp owns 32 accumulator rows (its TMEM lane quarter) and every second 32-column
    // chunk of the tile (4 chunks).  Chunks of consecutive tiles form one stream per warp;
    // without a blend operand the chunk buffers alternate and the load of chunk s+1 is in
    // flight while chunk s is computed; with a blend operand (K = 6144 launches, long MMAs)
    // buffer 0 holds the residual / result and buffer 1 the blend operand.
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    float* const buf0 = reinterpret_cast<float*>(epi_stage) + ew * 2 * (RT_CHUNK_BYTES / 4);
    auto bufp = [&](int b) { return buf0 + b * (RT_CHUNK_BYTES / 4); };
    uint64_t* rb = rt_bar + 2 * ew;
    const bool blend = p.blend_x != nullptr;
    const int rpi = static_cast<int>(p.rows_per_item);
    const int row_off = static_cast<int>(rank) * BM + quarter * 32;
    auto issue = [&](int tile, int k, int b) {     // lane 0 only
      const int m0 = (tile / n_blocks) * 2 * BM + row_off;
      const int n0 = (tile % n_blocks) * BNT + (half + 2 * k) * 32;
      mbar_expect_tx(&rb[b], blend ? 2 * RT_CHUNK_BYTES : RT_CHUNK_BYTES);
      tma_load_2d(&tmap_r, &rb[b], bufp(b), n0, m0, kEvictFirst);
      if (blend) tma_load_2d(&tmap_x, &rb[b], bufp(1), n0, m0, kEvictFirst);
    };
    // L2 prefetch of a tile's residual / blend rows (512 B per thread), one tile ahead of use
    auto prefetch_tile = [&](int tile) {
      if (tile < num_tiles)
        prefetch_resid_tile<EPI>(p, (tile / n_blocks) * 2 * BM + row_off + lane, M, (tile % n_blocks) * BNT, N, half, BNT / 2);
    };
    uint32_t seq = 0;
    prefetch_tile(cluster_id);
    if (cluster_id < num_tiles && lane == 0) issue(cluster_id, 0, 0);
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++it) {
      const int m_blk = tile / n_blocks;
      const int n_blk = tile % n_blocks;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      prefetch_tile(tile + n_clusters);
      const int m = m_blk * 2 * BM + row_off + lane;
      const int mc = m < M ? m : M - 1;            // tail rows: results are clipped by the store
      const int item = rpi > 0 ? mc / rpi : 0;
      const float alpha = blend ? __ldg(p.alpha + (p.rows_per_batch > 0 ? mc / static_cast<int>(p.rows_per_batch) : 0)) : 0.f;
      const uint32_t taddr = tmem_base + as * BNT + (static_cast<uint32_t>(quarter * 32) << 16);
      constexpr int NCH = BNT / 64;               // chunks per warp per tile
#pragma unroll 1
      for (int k = 0; k < NCH; ++k, ++seq) {
        const int cur = blend ? 0 : static_cast<int>(seq & 1);
        if (!blend && lane == 0) {                 // next chunk of the stream into the other buffer
          const bool next_here = k < NCH - 1;
          const int ntile = next_here ? tile : tile + n_clusters;
          if (ntile < num_tiles) {
            bulk_wait_read0();                     // the store that last read that buffer is done
            issue(ntile, next_here ? k + 1 : 0, cur ^ 1);
          }
        }
        const int n0 = n_blk * BNT + (half + 2 * k) * 32;
        const bool live = n0 < N;                  // chunks past N (N % 256 != 0) only keep the stream in step
        RtVecs vv;
        rt_load_vecs(vv, p, n0, item, live);
        if (k == 0) {
          mbar_wait(&tfull_bar[as], aphase);
          tc_fence_after();
        }
        uint32_t acc[32];
        tmem_ld32(taddr + (half + 2 * k) * 32, acc);
        mbar_wait(&rb[cur], blend ? (seq & 1) : ((seq >> 1) & 1));
        tmem_ld_wait();
        if (k == NCH - 1) {                        // accumulator fully read: hand TMEM back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&tempty_bar[as]), 0));
        }
        if (live) rt_chunk_math(acc, vv, bufp(cur), blend ? bufp(1) : nullptr, p, alpha, lane);
        fence_proxy_async();                       // generic-proxy writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) {
          if (live) {
            tma_store_2d(&tmap_o, bufp(cur), n0, m_blk * 2 * BM + row_off);
            bulk_commit();
          }
          if (blend) {                             // single-buffered: refill for the next chunk now
            const bool next_here = k < NCH - 1;
            const int ntile = next_here ? tile : tile + n_clusters;
            if (ntile < num_tiles) {
              bulk_wait_read0();
              issue(ntile, next_here ? k + 1 : 0, 0);
            }
          }
        }
      }
    }
    if (lane == 0) bulk_wait0();                   // results are in global memory before exit
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

int g_resid_tma = -1;   // -1: env DWM_RESID_TMA (default 1); dwm_b200_set_option("resid_tma", 0 | 1)

// The TMA residual path needs a plain [M, N] fp32 residual (row m -> row m), 16-byte pitches
// and whole 32-column chunks; the pos-embed (row modulo) and per-item residual forms stay on
// the register path.
static bool resid_tma_ok(const dwm_linear_args* a) {
  if (g_resid_tma < 0) {
    const char* e = getenv("DWM_RESID_TMA");
    g_resid_tma = (e && e[0] == '0') ? 0 : 1;
  }
  if (g_resid_tma != 1 || a->epilogue != DWM_EPI_RESID || !a->resid || a->resid_row_mod != 0) return false;
  if (a->N % 32 || a->ldr % 4 || a->ldo % 4 || a->ldr < a->N || a->ldo < a->N) return false;
  if ((reinterpret_cast<uintptr_t>(a->resid) & 15) || (reinterpret_cast<uintptr_t>(a->out) & 15)) return false;
  if (a->bias && (reinterpret_cast<uintptr_t>(a->bias) & 15)) return false;
  if (a->gate && ((reinterpret_cast<uintptr_t>(a->gate) & 15) || a->gate_ld % 4)) return false;
  if (a->blend_x && ((reinterpret_cast<uintptr_t>(a->blend_x) & 15) || a->ldx % 4 || a->ldx < a->N)) return false;
  return true;
}

int g_gemm_bn = 0;   // dwm_b200_set_option("gemm_bn", 0 = by wave efficiency | 128 | 256)

// Wave efficiency of the persistent schedule on the SM pairs: tiles / (waves * clusters).
static double wave_eff(long long tiles, int pairs) {
  const long long waves = (tiles + pairs - 1) / pairs;
  return static_cast<double>(tiles) / static_cast<double>(waves * pairs);
}
// 128-column tiles cost ~50 % more shared-memory operand traffic per FLOP and twice the
// per-tile bookkeeping: measured on B200 (tools/shard_emulate.py 8, r02) they LOSE 5 % at
// M = 10 752 (wave efficiency 0.85 -> 0.97: 183 vs 173 us for N = 1536, K = 6144) and WIN 8 % at
// M = 3 696 (0.61 -> 0.81: 85 vs 92 us), so they are used only where they repair the tail
// wave by more than 15 points.
static bool narrow_tiles_pay(const dwm_linear_args* a) {
  static bool env_read = false;
  if (!env_read) {               // DWM_GEMM_BN = 128 | 256 forces a tile width (measurement aid)
    env_read = true;
    const char* e = getenv("DWM_GEMM_BN");
    if (e && g_gemm_bn == 0) g_gemm_bn = atoi(e);
  }
  if (g_gemm_bn == 128) return true;
  if (g_gemm_bn == 256) return false;
  const int pairs = sm_count() / 2;
  const long long mb = (a->M + 2 * BM - 1) / (2 * BM);
  const double e256 = wave_eff(mb * ((a->N + 255) / 256), pairs);
  const double e128 = wave_eff(mb * ((a->N + 127) / 128), pairs);
  return e128 > e256 + 0.15;
}

template <typename T, int EPI, bool RT, int BNT>
static int launch_gemm2(const dwm_linear_args* a, cudaStream_t stream) {
  using Cfg = G2Cfg<RT, BNT>;
  CUtensorMap ta, tb, tr, to, tx;
  int rc = make_tmap_2d(&ta, a->A, a->M, a->K, a->lda, BM, BK, 2);
  if (rc) return rc;
  rc = make_tmap_2d(&tb, a->W, a->N, a->K, a->ldw, BNT / 2, BK, 2);
  if (rc) return rc;
  if (RT) {
    rc = make_tmap_2d(&tr, a->resid, a->M, a->N, a->ldr, 32, 32, 4);
    if (rc) return rc;
    rc = make_tmap_2d(&to, a->out, a->M, a->N, a->ldo, 32, 32, 4);
    if (rc) return rc;
    if (a->blend_x) {
      rc = make_tmap_2d(&tx, a->blend_x, a->M, a->N, a->ldx, 32, 32, 4);
      if (rc) return rc;
    } else {
      tx = tr;
    }
  } else {
    tr = ta; to = ta; tx = ta;   // unused by the kernel
  }
  EpiParams p;
  fill_epi_params(p, a);
  auto kern = gemm2_tcgen05_kernel<T, EPI, RT, BNT>;
  static bool attr_set = false;
  if (!attr_set) {
    DWM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const long long m_blocks = (a->M + 2 * BM - 1) / (2 * BM);
  const long long n_blocks = (a->N + BNT - 1) / BNT;
  const long long tiles = m_blocks * n_blocks;
  const int pairs = sm_count() / 2;
  const int clusters = static_cast<int>(tiles < pairs ? tiles : pairs);
  kern<<<2 * clusters, GEMM_THREADS, Cfg::kSmemBytes, stream>>>(
      ta, tb, tr, to, tx, static_cast<int>(a->M), static_cast<int>(a->N), static_cast<int>(a->K), p);
  DWM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
static int dispatch2(const dwm_linear_args* a, cudaStream_t s) {
  switch (a->epilogue) {
    case DWM_EPI_STORE: return launch_gemm2<T, DWM_EPI_STORE, false, 256>(a, s);
    case DWM_EPI_GEGLU: return launch_gemm2<T, DWM_EPI_GEGLU, false, 256>(a, s);
    case DWM_EPI_QKNORM: return launch_gemm2<T, DWM_EPI_QKNORM, false, 256>(a, s);
    case DWM_EPI_RESID:
      if (!resid_tma_ok(a)) return launch_gemm2<T, DWM_EPI_RESID, false, 256>(a, s);
      return narrow_tiles_pay(a) ? launch_gemm2<T, DWM_EPI_RESID, true, 128>(a, s)
                                 : launch_gemm2<T, DWM_EPI_RESID, true, 256>(a, s);
    case DWM_EPI_F32: return launch_gemm2<T, DWM_EPI_F32, false, 256>(a, s);
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
