// tcgen05 / TMEM flash attention for contiguous token groups (head_dim 64):
// the joint [sample ; context] attention of JointTransformerBlock (seq 602) and the
// dual attn2 (seq 448) — 90 % of the attention time of a CTSD-3.5 step.
//
// Persistent CTAs (one per SM) walk work items (group, head, 128-query tile):
//   warp 0     TMA producer: Q tile (double buffered) and K / V blocks of 128 keys
//              (3-stage ring) straight out of the fused q|k|v buffer, 128B swizzle
//   warp 1     MMA issuer: S = Q·K_blk^T (tcgen05.mma 128x128x16, fp32 S in TMEM, double
//              buffered) and O_blk = P·V_blk (128x64x16, P from swizzled smem, V as an
//              MN-major operand so no transpose is materialised)
//   warps 2-5  softmax: one thread per query row; running max / sum in registers,
//              P -> bf16/fp16 into shared memory for the second MMA, running output
//              accumulated in REGISTERS (O_blk is read back from TMEM per block), so the
//              rescale by exp2(m_old - m_new) never touches TMEM.
// Masked / gathered regroupings (cross-view row-wise, temporal) stay on the mma.sync
// kernel in attention.cu.
#include "common.cuh"
#include "../../include/dwm_b200.h"

namespace dwm {

constexpr int TC_BQ = 128;                 // queries per tile
constexpr int TC_BK = 128;                 // keys per block
constexpr int TC_HD = 64;
constexpr int TC_KV_STAGES = 3;
constexpr int TC_TILE_BYTES = 128 * 64 * 2;   // one 128-row x 64-col 16-bit tile
constexpr int TC_THREADS = 192;
constexpr int TC_SMEM_BYTES = 2 * TC_TILE_BYTES                  /* Q x2 */
                              + TC_KV_STAGES * 2 * TC_TILE_BYTES /* K,V ring */
                              + 2 * 2 * TC_TILE_BYTES            /* P x2 (two 64-key halves) */
                              + 1024 + 512;

struct AttnTcParams {
  int groups, heads, seq, q_tiles, n_kb;
  long long group_stride;   // rows per group in the qkv buffer
  int D;
  void* out; long long ldo; long long out_group_stride;
  int split; void* out2; long long ldo2;
  float scale_log2;
};

// MN-major (N contiguous) B operand written by TMA with 128B swizzle: rows = K index
// (keys), 128 B per row = 64 N elements.  8-row groups are 1024 B apart (SBO); LBO is the
// stride between 64-element N blocks (unused for N = 64).
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(TC_TILE_BYTES >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T>
__global__ void __launch_bounds__(TC_THREADS, 1)
    attn_tc_kernel(const __grid_constant__ CUtensorMap tmap, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sq = smem;                                        // [2][16 KB]
  uint8_t* skv = sq + 2 * TC_TILE_BYTES;                     // [stages][K 16 KB | V 16 KB]
  uint8_t* sp = skv + TC_KV_STAGES * 2 * TC_TILE_BYTES;      // [2][2][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sp + 4 * TC_TILE_BYTES);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = bars + 2;       // [2]
  uint64_t* kv_full = bars + 4;       // [3]
  uint64_t* kv_empty = bars + 7;      // [3]
  uint64_t* s_full = bars + 10;       // [2]
  uint64_t* s_empty = bars + 12;      // [2]
  uint64_t* p_full = bars + 14;       // [2]
  uint64_t* p_empty = bars + 16;      // [2]
  uint64_t* o_full = bars + 18;       // [2]
  uint64_t* o_empty = bars + 20;      // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.groups * p.heads * p.q_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);  mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);  mbar_init(&s_empty[i], 4);
      mbar_init(&p_full[i], 4);  mbar_init(&p_empty[i], 1);
      mbar_init(&o_full[i], 1);  mbar_init(&o_empty[i], 4);
    }
    for (int i = 0; i < TC_KV_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_s = tmem_base;           // 2 x 128 columns
  const uint32_t tmem_o = tmem_base + 256;     // 2 x 64 columns

  // work item -> (group, head, q tile); heads fastest so neighbouring CTAs share rows
  auto decode = [&](int w, int& g, int& h, int& qt) {
    h = w % p.heads;
    const int r = w / p.heads;
    qt = r % p.q_tiles;
    g = r / p.q_tiles;
  };

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int it = 0, kvs = 0;
      uint32_t kvph = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        int g, h, qt;
        decode(w, g, h, qt);
        const int qb = it & 1;
        const uint32_t qph = (it >> 1) & 1;
        const int row0 = static_cast<int>(g * p.group_stride);
        mbar_wait(&q_empty[qb], qph ^ 1);
        mbar_expect_tx(&q_full[qb], TC_TILE_BYTES);
        tma_load_2d(&tmap, &q_full[qb], sq + qb * TC_TILE_BYTES, h * TC_HD, row0 + qt * TC_BQ, kEvictFirst);
        for (int kb = 0; kb < p.n_kb; ++kb) {
          mbar_wait(&kv_empty[kvs], kvph ^ 1);
          mbar_expect_tx(&kv_full[kvs], 2 * TC_TILE_BYTES);
          uint8_t* st = skv + kvs * 2 * TC_TILE_BYTES;
          tma_load_2d(&tmap, &kv_full[kvs], st, p.D + h * TC_HD, row0 + kb * TC_BK, kEvictLast);
          tma_load_2d(&tmap, &kv_full[kvs], st + TC_TILE_BYTES, 2 * p.D + h * TC_HD, row0 + kb * TC_BK, kEvictLast);
          if (++kvs == TC_KV_STAGES) { kvs = 0; kvph ^= 1; }
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
      int blk = 0;  // global block counter (S / P / O double buffers)
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int qb = it & 1;
        const uint32_t qph = (it >> 1) & 1;
        mbar_wait(&q_full[qb], qph);
        tc_fence_after();
        const uint64_t dq = umma_desc_sw128(smem_u32(sq + qb * TC_TILE_BYTES));
        int pv_stage = 0;      // kv stage of the block whose PV is pending
        int pv_blk = 0;
        for (int kb = 0; kb <= p.n_kb; ++kb) {
          if (kb < p.n_kb) {
            // ---- S = Q K^T for block kb ----
            const int sb = blk & 1;
            const uint32_t sph = (blk >> 1) & 1;
            mbar_wait(&kv_full[kvs], kvph);
            mbar_wait(&s_empty[sb], sph ^ 1);
            tc_fence_after();
            const uint64_t dk = umma_desc_sw128(smem_u32(skv + kvs * 2 * TC_TILE_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_s + sb * 128, dq + 2 * k, dk + 2 * k, idesc_s, k ? 1u : 0u);
            umma_commit(&s_full[sb]);
            if (kb == p.n_kb - 1) umma_commit(&q_empty[qb]);   // Q tile fully consumed
          }
          if (kb > 0) {
            // ---- O_blk = P V for block kb-1 ----
            const int pb = pv_blk & 1;
            const uint32_t pph = (pv_blk >> 1) & 1;
            mbar_wait(&p_full[pb], pph);
            mbar_wait(&o_empty[pb], pph ^ 1);
            tc_fence_after();
            const uint32_t pa = smem_u32(sp + pb * 2 * TC_TILE_BYTES);
            const uint32_t va = smem_u32(skv + pv_stage * 2 * TC_TILE_BYTES + TC_TILE_BYTES);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t dp = umma_desc_sw128(pa + (ks >> 2) * TC_TILE_BYTES) + 2 * (ks & 3);
              const uint64_t dv = umma_desc_sw128_mn(va + ks * 2048);
              umma_f16(tmem_o + pb * 64, dp, dv, idesc_o, ks ? 1u : 0u);
            }
            umma_commit(&o_full[pb]);
            umma_commit(&p_empty[pb]);
            umma_commit(&kv_empty[pv_stage]);
          }
          if (kb < p.n_kb) {
            pv_stage = kvs;
            pv_blk = blk;
            ++blk;
            if (++kvs == TC_KV_STAGES) { kvs = 0; kvph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ================= softmax / output (warps 2..5, thread = query row) =================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    int blk = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      int g, h, qt;
      decode(w, g, h, qt);
      float m = -INFINITY, l = 0.f;
      float acc[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] = 0.f;
      for (int kb = 0; kb < p.n_kb; ++kb, ++blk) {
        const int sb = blk & 1;
        const uint32_t sph = (blk >> 1) & 1;
        const int kvalid = p.seq - kb * TC_BK;   // keys < kvalid are real
        mbar_wait(&s_full[sb], sph);
        tc_fence_after();
        const uint32_t ts = tmem_s + sb * 128 + lane_off;
        const bool full = kvalid >= TC_BK;   // warp-uniform: only the last block is ragged
        // the whole 128-key score row of this thread in registers: 4 TMEM loads in flight,
        // one wait (the loads are NOT re-issued for the exp pass)
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
              if (c * 32 + j < kvalid) mx = fmaxf(mx, __uint_as_float(sr[c][j]));
        }
        // S buffer is free for the next QK^T as soon as it sits in registers
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[sb]);
        const float m_new = fmaxf(m, mx * p.scale_log2);
        const float corr = ex2_approx(m - m_new);
        m = m_new;
        // p = exp2(s*scale - m), written as 16-bit into the swizzled K-major P tiles
        mbar_wait(&p_empty[sb], sph ^ 1);
        uint8_t* pbase = sp + sb * 2 * TC_TILE_BYTES;
        float rs0 = 0.f, rs1 = 0.f;
        const float sc = p.scale_log2;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          // keys c*32 .. c*32+31 -> half (c>>1), 16-byte chunks (c&1)*4 .. +3 of this row
          uint8_t* dst = pbase + (c >> 1) * TC_TILE_BYTES + row * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = q * 8 + 2 * e;
              float p0 = ex2_approx(fmaf(__uint_as_float(sr[c][j]), sc, -m_new));
              float p1 = ex2_approx(fmaf(__uint_as_float(sr[c][j + 1]), sc, -m_new));
              if (!full) {
                if (c * 32 + j >= kvalid) p0 = 0.f;
                if (c * 32 + j + 1 >= kvalid) p1 = 0.f;
              }
              rs0 += p0;
              rs1 += p1;
              pk[e] = Cvt<T>::pack2(p0, p1);
            }
            const int chunk = (c & 1) * 4 + q;
            *reinterpret_cast<uint4*>(dst + ((chunk ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        const float rs = rs0 + rs1;
        l = l * corr + rs;
        tc_fence_before();
        fence_proxy_async();          // generic-proxy smem writes -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[sb]);
        // fold in the previous block's P·V, then rescale for this block's max
        if (kb > 0) {
          const int ob = (blk - 1) & 1;
          const uint32_t oph = ((blk - 1) >> 1) & 1;
          mbar_wait(&o_full[ob], oph);
          tc_fence_after();
          const uint32_t to = tmem_o + ob * 64 + lane_off;
          {
            uint32_t r0[32], r1[32];
            tmem_ld32(to, r0);
            tmem_ld32(to + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              acc[j] += __uint_as_float(r0[j]);
              acc[32 + j] += __uint_as_float(r1[j]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_empty[ob]);
        }
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] *= corr;
      }
      // last block's P·V
      {
        const int ob = (blk - 1) & 1;
        const uint32_t oph = ((blk - 1) >> 1) & 1;
        mbar_wait(&o_full[ob], oph);
        tc_fence_after();
        const uint32_t to = tmem_o + ob * 64 + lane_off;
        {
          uint32_t r0[32], r1[32];
          tmem_ld32(to, r0);
          tmem_ld32(to + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            acc[j] += __uint_as_float(r0[j]);
            acc[32 + j] += __uint_as_float(r1[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[ob]);
      }
      const int j = qt * TC_BQ + row;
      if (j < p.seq) {
        const float inv = 1.0f / l;
        T* dst;
        if (p.split > 0 && j >= p.split)
          dst = reinterpret_cast<T*>(p.out2) + (static_cast<long long>(g) * (p.seq - p.split) + (j - p.split)) * p.ldo2;
        else
          dst = reinterpret_cast<T*>(p.out) + (static_cast<long long>(g) * p.out_group_stride + j) * p.ldo;
        dst += h * TC_HD;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint4 v;
          v.x = Cvt<T>::pack2(acc[8 * q] * inv, acc[8 * q + 1] * inv);
          v.y = Cvt<T>::pack2(acc[8 * q + 2] * inv, acc[8 * q + 3] * inv);
          v.z = Cvt<T>::pack2(acc[8 * q + 4] * inv, acc[8 * q + 5] * inv);
          v.w = Cvt<T>::pack2(acc[8 * q + 6] * inv, acc[8 * q + 7] * inv);
          *reinterpret_cast<uint4*>(dst + 8 * q) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <typename T>
static int launch_attn_tc(const dwm_attention_args* a, cudaStream_t s) {
  const long long groups = a->group_dims[0];
  const long long rows_total = groups * a->group_strides[0];
  CUtensorMap tm;
  int rc = make_tmap_2d(&tm, a->qkv, rows_total, 3 * a->D, a->ld, 128, 64, 2);
  if (rc) return rc;
  AttnTcParams p;
  p.groups = static_cast<int>(groups);
  p.heads = a->heads;
  p.seq = a->seq;
  p.q_tiles = (a->seq + TC_BQ - 1) / TC_BQ;
  p.n_kb = (a->seq + TC_BK - 1) / TC_BK;
  p.group_stride = a->group_strides[0];
  p.D = static_cast<int>(a->D);
  p.out = a->out; p.ldo = a->ldo; p.out_group_stride = a->out_group_strides[0];
  p.split = a->split; p.out2 = a->out2; p.ldo2 = a->ldo2;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  auto kern = attn_tc_kernel<T>;
  static bool attr_set = false;
  if (!attr_set) {
    DWM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    attr_set = true;
  }
  const long long items = groups * a->heads * p.q_tiles;
  const int sms = sm_count();
  const int grid = static_cast<int>(items < sms ? items : sms);
  kern<<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(tm, p);
  DWM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Eligibility: one contiguous run of rows per group, same buffer for q and k/v, no mask.
bool attn_tc_eligible(const dwm_attention_args* a) {
  return a->kv == nullptr && a->mask == nullptr && a->group_dims[1] == 1 && a->group_dims[2] == 1 &&
         a->inner == a->seq && a->stride_inner == 1 && a->out_stride_inner == 1 && a->seq > 64 &&
         a->group_strides[0] >= a->seq && a->group_dims[0] * a->group_strides[0] < (1ll << 31);
}

int attn_tc_launch(const dwm_attention_args* a, cudaStream_t s) {
  if (a->dtype == DWM_BF16) return launch_attn_tc<__nv_bfloat16>(a, s);
  if (a->dtype == DWM_F16) return launch_attn_tc<__half>(a, s);
  set_last_error("dwm_b200_attention: dtype must be DWM_BF16 or DWM_F16, got %d", a->dtype);
  return -1;
}

}  // namespace dwm
