// tcgen05 / TMEM flash attention, second generation: TWO co-resident CTAs per SM and the
// running output kept in TMEM.
//
// A one-CTA-per-SM design (the first-generation kernel of round 1) is bound by its 4 softmax
// warps (one per SM sub-partition): every score costs one MUFU ex2 plus ~4 issue slots, and a
// single warp per sub-partition cannot keep the MUFU pipe full while it also moves P and folds
// P·V into a register accumulator.
// Here each CTA is slimmed down so that two fit on one SM (112 KB smem, 256 TMEM columns,
// 128 registers/thread at launch, re-partitioned 40 / 216 with setmaxnreg): the sub-partitions see two softmax warps each, and one CTA's
// tensor-core work overlaps the other's exponentials.
//
//   warp 0     TMA producer: Q tile (single buffer) and K / V blocks of 128 keys (2 stages)
//   warp 1     MMA issuer: S = Q·K_blk^T (128x128x16, one S buffer — the softmax threads
//              pull the row into registers and release it at once) and O += P·V_blk
//              accumulated IN TMEM over the whole key sequence
//   warps 2-3  idle (they only exist so that the softmax warps form an aligned warpgroup
//              for setmaxnreg)
//   warps 4-7  softmax, thread = query row.  O is rescaled only when a row's running max
//              grew by more than 2^8 since the last rescale ("lazy rescale"): P is computed
//              against the stale reference max m_ref (values <= 256, exact in fp32 sums and
//              representable in bf16 / fp16), so the common case touches neither O nor l.
//              out = O / l is independent of m_ref.
//
// Gathered sequences (template flag G): cross-view row-wise attention
// "(bt v) (h w) c -> (bt h) (v w) c" with its [B,V,V] view mask, and temporal row-wise
// attention "(b t v) (h w) c -> (b v h) (t w) c" run on the same pipeline.  A sequence is
// n_out "outer" units (views / frames) of `inner` contiguous tokens (one latent row); a 5-D
// tensor map (col, w, outer, g1, g0) lets ONE bulk tensor load fetch a tile of `upt` whole
// units (upt * inner <= 128 rows) straight from the un-permuted q|k|v buffer — the
// reference's two 264 MB permutes and its [512,168,168] mask never exist.  The view mask is
// a per-row bit set over key units, expanded once per key block to a 128-bit column mask.
#include <string.h>

#include "common.cuh"
#include "../../include/dwm_b200.h"

namespace dwm {

namespace tc2 {
constexpr int BQ = 128, BK = 128, HD = 64;
constexpr int KV_STAGES = 2;
constexpr int TILE = 128 * 64 * 2;            // one 128-row x 64-col 16-bit tile
constexpr int THREADS = 256;               // warpgroup 0: TMA, MMA, 2 idle warps; warpgroup 1: softmax
constexpr int BAR_BYTES = 256;                // barriers + TMEM pointer at the start of dynamic smem
constexpr int TILES_BYTES = TILE /*Q*/ + KV_STAGES * 2 * TILE /*K,V*/ + 2 * TILE /*P: two 64-key halves*/;
constexpr int SMEM_BYTES = 1024 + TILES_BYTES;   // 115 712 B: two CTAs + 2 x 1 KB reserved = 228 KB
constexpr float RESCALE_LOG2 = 8.0f;
}  // namespace tc2

struct AttnTc2Params {
  int groups, heads, seq, q_tiles, n_kb;
  long long group_stride;
  int D;
  void* out; long long ldo; long long out_group_stride;
  int split; void* out2; long long ldo2;
  float scale_log2;
  // gathered sequences (G): group g = g0 * g1n + i1
  int inner, n_out, upt, g1n;
  long long out_gs1, out_so;
  const unsigned char* mask; int mask_div, mask_n;
};

__device__ __forceinline__ uint64_t umma_desc_sw128_mn2(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(tc2::TILE >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ float ex2_approx2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T, bool G>
__global__ void __launch_bounds__(tc2::THREADS, 2)
    attn_tc2_kernel(const __grid_constant__ CUtensorMap tmap, const AttnTc2Params p) {
  using namespace tc2;
  extern __shared__ uint8_t smem_raw[];
  // [barriers 256 B][pad][tiles, 1024-aligned]; with a 1024-aligned base the tiles start at
  // +1024 and end exactly at SMEM_BYTES
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t tiles_off = ((base + BAR_BYTES + 1023u) & ~1023u) - base;
  if (tiles_off + TILES_BYTES > SMEM_BYTES) {
    if (threadIdx.x == 0) printf("attn_tc2_kernel: dynamic smem base %u not 1 KB aligned enough\n", base);
    __trap();
  }
  uint8_t* sq = smem_raw + tiles_off;                      // Q  [16 KB]
  uint8_t* skv = sq + TILE;                                // [stages][K 16 KB | V 16 KB]
  uint8_t* sp = skv + KV_STAGES * 2 * TILE;                // P  [2 x 16 KB]
  uint64_t* q_full = bars;                // [1]
  uint64_t* q_empty = bars + 1;           // [1]
  uint64_t* kv_full = bars + 2;           // [2]
  uint64_t* kv_empty = bars + 4;          // [2]
  uint64_t* s_full = bars + 6;
  uint64_t* s_empty = bars + 7;
  uint64_t* p_full = bars + 8;
  uint64_t* p_empty = bars + 9;
  uint64_t* o_full = bars + 10;           // one completion per P·V (block granularity)
  uint64_t* o_empty = bars + 11;          // one completion per work item
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.groups * p.heads * p.q_tiles;
  // G: a tile holds upt * inner <= 128 rows; the rows behind them are never written by TMA and
  // must be finite (P = 0 times a stale NaN in V would poison O): zero all tiles once
  const uint32_t tile_tx = G ? static_cast<uint32_t>(p.upt * p.inner * 128) : static_cast<uint32_t>(TILE);
  if constexpr (G) {
    uint4* z = reinterpret_cast<uint4*>(sq);
    for (int i = threadIdx.x; i < TILES_BYTES / 16; i += THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(q_full, 1);  mbar_init(q_empty, 1);
    mbar_init(s_full, 1);  mbar_init(s_empty, 4);
    mbar_init(p_full, 4);  mbar_init(p_empty, 1);
    mbar_init(o_full, 1);  mbar_init(o_empty, 4);
    for (int i = 0; i < KV_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_s = tmem_base;           // 128 columns
  const uint32_t tmem_o = tmem_base + 128;     // 64 columns

  auto decode = [&](int w, int& g, int& h, int& qt) {
    h = w % p.heads;
    const int r = w / p.heads;
    qt = r % p.q_tiles;
    g = r / p.q_tiles;
  };

  // register re-partition (setmaxnreg needs aligned warpgroups): 2 CTAs x 256 threads start
  // at 128 registers; the control warpgroup shrinks to 40, the softmax warpgroup grows to
  // 216 so that the 128-column score row stays in registers without spills
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int it = 0, kvs = 0;
      uint32_t kvph = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        int g, h, qt;
        decode(w, g, h, qt);
        const int row0 = static_cast<int>(g * p.group_stride);
        const int g0 = G ? g / p.g1n : 0, i1 = G ? g - g0 * p.g1n : 0;
        // the NEXT item's Q / K / V boxes start their trip from HBM to L2 now: its loads can
        // only be issued once this item's P·V MMAs have released the K,V stages, and with 2..5
        // key blocks per item that latency would otherwise be exposed once per item
        if (w + static_cast<int>(gridDim.x) < n_items) {
          int g2, h2, qt2;
          decode(w + gridDim.x, g2, h2, qt2);
          if constexpr (G) {
            const int g02 = g2 / p.g1n, i12 = g2 - g02 * p.g1n;
            tma_prefetch_5d(&tmap, h2 * HD, 0, qt2 * p.upt, i12, g02);
            for (int kb = 0; kb < p.n_kb; ++kb) {
              tma_prefetch_5d(&tmap, p.D + h2 * HD, 0, kb * p.upt, i12, g02);
              tma_prefetch_5d(&tmap, 2 * p.D + h2 * HD, 0, kb * p.upt, i12, g02);
            }
          } else {
            const int r2 = static_cast<int>(g2 * p.group_stride);
            tma_prefetch_2d(&tmap, h2 * HD, r2 + qt2 * BQ);
            for (int kb = 0; kb < p.n_kb; ++kb) {
              tma_prefetch_2d(&tmap, p.D + h2 * HD, r2 + kb * BK);
              tma_prefetch_2d(&tmap, 2 * p.D + h2 * HD, r2 + kb * BK);
            }
          }
        }
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_expect_tx(q_full, tile_tx);
        if constexpr (G) tma_load_5d(&tmap, q_full, sq, h * HD, 0, qt * p.upt, i1, g0, kEvictFirst);
        else tma_load_2d(&tmap, q_full, sq, h * HD, row0 + qt * BQ, kEvictFirst);
        for (int kb = 0; kb < p.n_kb; ++kb) {
          mbar_wait(&kv_empty[kvs], kvph ^ 1);
          mbar_expect_tx(&kv_full[kvs], 2 * tile_tx);
          uint8_t* st = skv + kvs * 2 * TILE;
          if constexpr (G) {
            tma_load_5d(&tmap, &kv_full[kvs], st, p.D + h * HD, 0, kb * p.upt, i1, g0, kEvictLast);
            tma_load_5d(&tmap, &kv_full[kvs], st + TILE, 2 * p.D + h * HD, 0, kb * p.upt, i1, g0, kEvictLast);
          } else {
            tma_load_2d(&tmap, &kv_full[kvs], st, p.D + h * HD, row0 + kb * BK, kEvictLast);
            tma_load_2d(&tmap, &kv_full[kvs], st + TILE, 2 * p.D + h * HD, row0 + kb * BK, kEvictLast);
          }
          if (++kvs == KV_STAGES) { kvs = 0; kvph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc(128, 128, Cvt<T>::kUmmaFmt);
      constexpr uint32_t idesc_o = umma_idesc(128, 64, Cvt<T>::kUmmaFmt) | (1u << 16);  // B MN-major
      int it = 0, kvs = 0;
      uint32_t kvph = 0;
      int blk = 0;  // global key-block counter: phases of s_*, p_*, o_full
      const uint64_t dq = umma_desc_sw128(smem_u32(sq));
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        mbar_wait(q_full, it & 1);
        tc_fence_after();
        int pv_stage = 0, pv_blk = 0;
        for (int kb = 0; kb <= p.n_kb; ++kb) {
          if (kb < p.n_kb) {
            // ---- S = Q K^T for block kb ----
            mbar_wait(&kv_full[kvs], kvph);
            mbar_wait(s_empty, (blk & 1) ^ 1);
            tc_fence_after();
            const uint64_t dk = umma_desc_sw128(smem_u32(skv + kvs * 2 * TILE));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k ? 1u : 0u);
            umma_commit(s_full);
            if (kb == p.n_kb - 1) umma_commit(q_empty);   // Q tile fully consumed
          }
          if (kb > 0) {
            // ---- O (+)= P V for block kb-1 ----
            mbar_wait(p_full, pv_blk & 1);
            if (kb == 1) mbar_wait(o_empty, (it & 1) ^ 1);   // previous item's O has been read out
            tc_fence_after();
            const uint32_t pa = smem_u32(sp);
            const uint32_t va = smem_u32(skv + pv_stage * 2 * TILE + TILE);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t dp = umma_desc_sw128(pa + (ks >> 2) * TILE) + 2 * (ks & 3);
              const uint64_t dv = umma_desc_sw128_mn2(va + ks * 2048);
              umma_f16(tmem_o, dp, dv, idesc_o, (kb > 1 || ks) ? 1u : 0u);
            }
            umma_commit(o_full);
            umma_commit(p_empty);
            umma_commit(&kv_empty[pv_stage]);
          }
          if (kb < p.n_kb) {
            pv_stage = kvs;
            pv_blk = blk;
            ++blk;
            if (++kvs == KV_STAGES) { kvs = 0; kvph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ================= softmax / output (warps 4..7, thread = query row) =================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t ts = tmem_s + lane_off;
    const uint32_t to = tmem_o + lane_off;
    const float sc = p.scale_log2;
    int blk = 0, it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      int g, h, qt;
      decode(w, g, h, qt);
      float m_ref = -INFINITY, l = 0.f;
      // G: this thread's query token = (outer unit qo, token wi of the unit); `allowed` = key
      // units it may attend to (the [B,V,V] mask row; everything for padding rows)
      int qo = 0, wi = 0;
      bool row_ok = true;
      uint32_t allowed = 0xffffffffu;
      if constexpr (G) {
        const int u = row / p.inner;
        wi = row - u * p.inner;
        qo = qt * p.upt + u;
        row_ok = u < p.upt && qo < p.n_out;
        if (p.mask && row_ok) {
          const int g0 = g / p.g1n;
          const unsigned char* mr = p.mask + (static_cast<long long>(g0 / p.mask_div) * p.mask_n + qo) * p.mask_n;
          allowed = 0u;
          for (int ko = 0; ko < p.n_out; ++ko) allowed |= (__ldg(mr + ko) != 0 ? 1u : 0u) << ko;
        }
      }
      for (int kb = 0; kb < p.n_kb; ++kb, ++blk) {
        const uint32_t ph = blk & 1;
        int kvalid = p.seq - kb * BK;   // keys < kvalid are real
        uint32_t cmw[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        if constexpr (G) {
          // 128-bit column mask of this key block: unit u2 covers columns [u2*inner, (u2+1)*inner)
          unsigned __int128 cm = 0;
          const unsigned __int128 ones = p.inner >= 128 ? ~static_cast<unsigned __int128>(0)
                                                        : ((static_cast<unsigned __int128>(1) << p.inner) - 1);
          for (int u2 = 0; u2 < p.upt; ++u2) {
            const int ko = kb * p.upt + u2;
            if (ko < p.n_out && ((allowed >> ko) & 1u)) cm |= ones << (u2 * p.inner);
          }
          cmw[0] = static_cast<uint32_t>(cm);
          cmw[1] = static_cast<uint32_t>(cm >> 32);
          cmw[2] = static_cast<uint32_t>(cm >> 64);
          cmw[3] = static_cast<uint32_t>(cm >> 96);
          kvalid = 0;                    // G always takes the masked ("ragged") path
        }
        mbar_wait(s_full, ph);
        tc_fence_after();
        const bool full = kvalid >= BK;       // warp-uniform: only the last block is ragged
        uint32_t sr[4][32];
        tmem_ld32(ts, sr[0]);
        tmem_ld32(ts + 32, sr[1]);
        tmem_ld32(ts + 64, sr[2]);
        tmem_ld32(ts + 96, sr[3]);
        tmem_ld_wait();
        float mx = -INFINITY;
        if (full) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j = 0; j < 32; j += 2)
              mx = fmaxf(mx, fmaxf(__uint_as_float(sr[c][j]), __uint_as_float(sr[c][j + 1])));
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (G ? ((cmw[c] >> j) & 1u) != 0u : (c * 32 + j < kvalid)) mx = fmaxf(mx, __uint_as_float(sr[c][j]));
        }
        // the S buffer is free for the next QK^T as soon as the row sits in registers
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty);
        const float m_new = fmaxf(m_ref, mx * sc);
        // lazy rescale: warp-uniform decision, per-row factor
        if (__any_sync(0xffffffffu, m_new - m_ref > RESCALE_LOG2)) {
          // 0 on the first block (m_ref = -inf); a row whose keys were all masked so far
          // (m_new = -inf, G only) keeps l = 0 / O = 0 instead of picking up inf - inf
          const float corr = m_new == -INFINITY ? 1.0f : ex2_approx2(m_ref - m_new);
          l *= corr;
          m_ref = m_new;
          if (kb > 0) {
            // O holds blocks 0..kb-1 once P·V(kb-1) has completed; P·V(kb) cannot start
            // before this warp arrives on p_full below, so o_full is at most one phase ahead
            mbar_wait(o_full, (blk - 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {     // 32 columns at a time: register pressure
              uint32_t r[32];
              tmem_ld32(to + 32 * half, r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * corr);
              tmem_st32(to + 32 * half, r);
            }
            tmem_st_wait();
          }
        }
        // p = exp2(s*scale - m_ref), written as 16-bit into the swizzled K-major P tiles
        mbar_wait(p_empty, ph ^ 1);
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint8_t* dst = sp + (c >> 1) * TILE + row * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = q * 8 + 2 * e;
              float p0 = ex2_approx2(fmaf(__uint_as_float(sr[c][j]), sc, -m_ref));
              float p1 = ex2_approx2(fmaf(__uint_as_float(sr[c][j + 1]), sc, -m_ref));
              if (!full) {
                if (G ? ((cmw[c] >> j) & 1u) == 0u : (c * 32 + j >= kvalid)) p0 = 0.f;
                if (G ? ((cmw[c] >> (j + 1)) & 1u) == 0u : (c * 32 + j + 1 >= kvalid)) p1 = 0.f;
              }
              rs0 += p0;
              rs1 += p1;
              pk[e] = Cvt<T>::pack2(p0, p1);
            }
            const int chunk = (c & 1) * 4 + q;
            *reinterpret_cast<uint4*>(dst + ((chunk ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        l += rs0 + rs1;
        tc_fence_before();
        fence_proxy_async();          // generic-proxy smem writes -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      // all P·V of this item done -> O / l
      mbar_wait(o_full, (blk - 1) & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld32(to, r0);
      tmem_ld32(to + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      const int j = qt * BQ + row;
      if (G ? row_ok : (j < p.seq)) {
        const float inv = 1.0f / l;
        T* dst;
        if constexpr (G) {
          const int g0 = g / p.g1n, i1 = g - g0 * p.g1n;
          dst = reinterpret_cast<T*>(p.out) +
                (static_cast<long long>(g0) * p.out_group_stride + static_cast<long long>(i1) * p.out_gs1 +
                 static_cast<long long>(qo) * p.out_so + wi) * p.ldo;
        } else if (p.split > 0 && j >= p.split) {
          dst = reinterpret_cast<T*>(p.out2) + (static_cast<long long>(g) * (p.seq - p.split) + (j - p.split)) * p.ldo2;
        } else {
          dst = reinterpret_cast<T*>(p.out) + (static_cast<long long>(g) * p.out_group_stride + j) * p.ldo;
        }
        dst += h * HD;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v;
          v.x = Cvt<T>::pack2(__uint_as_float(r0[8 * q]) * inv, __uint_as_float(r0[8 * q + 1]) * inv);
          v.y = Cvt<T>::pack2(__uint_as_float(r0[8 * q + 2]) * inv, __uint_as_float(r0[8 * q + 3]) * inv);
          v.z = Cvt<T>::pack2(__uint_as_float(r0[8 * q + 4]) * inv, __uint_as_float(r0[8 * q + 5]) * inv);
          v.w = Cvt<T>::pack2(__uint_as_float(r0[8 * q + 6]) * inv, __uint_as_float(r0[8 * q + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + 8 * q) = v;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v;
          v.x = Cvt<T>::pack2(__uint_as_float(r1[8 * q]) * inv, __uint_as_float(r1[8 * q + 1]) * inv);
          v.y = Cvt<T>::pack2(__uint_as_float(r1[8 * q + 2]) * inv, __uint_as_float(r1[8 * q + 3]) * inv);
          v.z = Cvt<T>::pack2(__uint_as_float(r1[8 * q + 4]) * inv, __uint_as_float(r1[8 * q + 5]) * inv);
          v.w = Cvt<T>::pack2(__uint_as_float(r1[8 * q + 6]) * inv, __uint_as_float(r1[8 * q + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + 32 + 8 * q) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <typename T, bool G>
static int launch_attn_tc2(const dwm_attention_args* a, cudaStream_t s) {
  using namespace tc2;
  AttnTc2Params p;
  memset(&p, 0, sizeof(p));
  CUtensorMap tm;
  long long groups;
  if (G) {
    // merge the (optional) third group dim into the second: row offset i1*gs1 + i2*gs2 with
    // gs1 == gd2*gs2 is (i1*gd2 + i2)*gs2 (checked by attn_tcg_eligible)
    const long long gd1 = a->group_dims[1] * a->group_dims[2];
    const long long gs1 = a->group_dims[2] > 1 ? a->group_strides[2] : a->group_strides[1];
    const long long ogs1 = a->group_dims[2] > 1 ? a->out_group_strides[2] : a->out_group_strides[1];
    p.inner = a->inner;
    p.n_out = a->seq / a->inner;
    p.upt = 128 / a->inner;
    if (p.upt > p.n_out) p.upt = p.n_out;
    p.g1n = static_cast<int>(gd1);
    p.out_gs1 = ogs1;
    p.out_so = a->out_stride_outer;
    p.mask = a->mask; p.mask_div = a->mask_div; p.mask_n = a->n_outer;
    groups = a->group_dims[0] * gd1;
    const uint64_t eb = 2;
    const uint64_t dims[5] = {static_cast<uint64_t>(3 * a->D), static_cast<uint64_t>(a->inner),
                              static_cast<uint64_t>(p.n_out), static_cast<uint64_t>(gd1),
                              static_cast<uint64_t>(a->group_dims[0])};
    const uint64_t st[4] = {static_cast<uint64_t>(a->ld) * eb, static_cast<uint64_t>(a->stride_outer * a->ld) * eb,
                            static_cast<uint64_t>(gs1 * a->ld) * eb,
                            static_cast<uint64_t>(a->group_strides[0] * a->ld) * eb};
    const uint32_t box[5] = {64u, static_cast<uint32_t>(a->inner), static_cast<uint32_t>(p.upt), 1u, 1u};
    int rc = make_tmap_nd(&tm, a->qkv, 5, dims, st, box, 2);
    if (rc) return rc;
    p.q_tiles = (p.n_out + p.upt - 1) / p.upt;
    p.n_kb = p.q_tiles;
  } else {
    groups = a->group_dims[0];
    const long long rows_total = groups * a->group_strides[0];
    int rc = make_tmap_2d(&tm, a->qkv, rows_total, 3 * a->D, a->ld, 128, 64, 2);
    if (rc) return rc;
    p.q_tiles = (a->seq + BQ - 1) / BQ;
    p.n_kb = (a->seq + BK - 1) / BK;
  }
  p.groups = static_cast<int>(groups);
  p.heads = a->heads;
  p.seq = a->seq;
  p.group_stride = a->group_strides[0];
  p.D = static_cast<int>(a->D);
  p.out = a->out; p.ldo = a->ldo; p.out_group_stride = a->out_group_strides[0];
  p.split = a->split; p.out2 = a->out2; p.ldo2 = a->ldo2;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  auto kern = attn_tc2_kernel<T, G>;
  static bool attr_set = false;
  if (!attr_set) {
    DWM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    DWM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  const long long items = groups * a->heads * p.q_tiles;
  const long long slots = 2ll * sm_count();
  const int grid = static_cast<int>(items < slots ? items : slots);
  kern<<<grid, THREADS, SMEM_BYTES, s>>>(tm, p);
  DWM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// contiguous, unmasked head_dim-64 sequences (joint / dual attention, UNet spatial attention)
bool attn_tc_eligible(const dwm_attention_args* a) {
  return a->kv == nullptr && a->mask == nullptr && a->group_dims[1] == 1 && a->group_dims[2] == 1 &&
         a->inner == a->seq && a->stride_inner == 1 && a->out_stride_inner == 1 && a->seq > 64 &&
         a->group_strides[0] >= a->seq && a->group_dims[0] * a->group_strides[0] < (1ll << 31);
}

// gathered sequences of whole units of `inner` contiguous tokens (cross-view / temporal
// row-wise), optional [B, n_outer, n_outer] unit mask
bool attn_tcg_eligible(const dwm_attention_args* a) {
  if (a->kv != nullptr || a->split > 0 || a->seq <= 64 || a->inner <= 0 || a->inner > 128) return false;
  if (a->inner == a->seq && a->mask == nullptr) return false;      // contiguous: the 2-D path
  if (a->seq % a->inner || a->stride_inner != 1 || a->out_stride_inner != 1) return false;
  const long long n_out = a->seq / a->inner;
  if (n_out > 32 || n_out > 256) return false;
  if (a->mask && a->n_outer != n_out) return false;
  if (a->group_dims[2] > 1 &&
      (a->group_strides[1] != a->group_dims[2] * a->group_strides[2] ||
       a->out_group_strides[1] != a->group_dims[2] * a->out_group_strides[2]))
    return false;
  if (a->stride_outer <= 0 || a->group_strides[0] <= 0) return false;
  const long long groups = a->group_dims[0] * a->group_dims[1] * a->group_dims[2];
  return groups * a->heads * 8 < (1ll << 31);
}

int attn_tc2_launch(const dwm_attention_args* a, cudaStream_t s) {
  if (a->dtype == DWM_BF16) return launch_attn_tc2<__nv_bfloat16, false>(a, s);
  if (a->dtype == DWM_F16) return launch_attn_tc2<__half, false>(a, s);
  set_last_error("dwm_b200_attention: dtype must be DWM_BF16 or DWM_F16, got %d", a->dtype);
  return -1;
}

int attn_tcg_launch(const dwm_attention_args* a, cudaStream_t s) {
  if (a->dtype == DWM_BF16) return launch_attn_tc2<__nv_bfloat16, true>(a, s);
  if (a->dtype == DWM_F16) return launch_attn_tc2<__half, true>(a, s);
  set_last_error("dwm_b200_attention: dtype must be DWM_BF16 or DWM_F16, got %d", a->dtype);
  return -1;
}

}  // namespace dwm
