// Fused-epilogue machinery shared by the 1-CTA and 2-CTA tcgen05 GEMM kernels.
#pragma once
#include "common.cuh"
#include "../../include/dwm_b200.h"

namespace dwm {

constexpr int BM = 128;   // accumulator rows per CTA (TMEM lanes)
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;  // two warps per TMEM lane quarter, interleaved over column chunks
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int EPI_STAGE_BYTES = EPI_WARPS * 32 * 32 * 4;  // per epilogue warp: 32x32 fp32

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
  long long ldr, resid_row_mod;   // resid_row_mod < 0: one residual row per ITEM (m / rows_per_item)
  const float* gate;
  long long gate_ld;
  const float* blend_x;
  long long ldx;
  const float* alpha;
  long long rows_per_batch;
  void* peer_out[8];
  int n_peers;
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

// RESID arithmetic, spelled with explicit roundings so that every kernel variant (1-CTA /
// 2-CTA, register / TMA-staged epilogue, convolution) produces the same bits whatever the
// compiler would contract: v = fma(acc + bias, gate, resid); blend: fma(alpha, x, (1-alpha) v).
// Absent bias / gate are 0 / 1, which leaves the value exact.
__device__ __forceinline__ float resid_elem(float acc, float b, float g, float r) {
  return __fmaf_rn(__fadd_rn(acc, b), g, r);
}
__device__ __forceinline__ float blend_elem(float a, float a1, float x, float v) {
  return __fmaf_rn(a, x, __fmul_rn(a1, v));
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

// Row mapping of an accumulator tile.  Linear (GEMM): tile row r is global row m_base + r,
// valid while < M.  Pixel tile (implicit-GEMM convolution): the 128 accumulator rows are a
// bw x bh patch of an image of width img_w; row r = (r / bw, r % bw) maps to global row
// m_base + (r / bw) * img_w + r % bw and is valid inside the patch / image limits.
struct TileGeom {
  int bw;      // 0 => linear
  int rows;    // bw * bh
  int w_lim;   // img_w - w0
  int h_lim;   // img_h - h0
  int img_w;
  int tile_cols;   // accumulator columns of this tile (0 => BN)
};

template <typename T, int EPI>
__device__ __forceinline__ void drain_tile(uint32_t taddr, float4* stg, int m_base, int row0, int M,
                                           int n_tile0, int N, const EpiParams& p, int lane,
                                           int half, const TileGeom geom = TileGeom{0, 0, 0, 0, 0, 0}) {
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
    int m;
    bool valid;
    if (geom.bw > 0) {
      const int r = row0 + it * 4 + rs;
      const int ph = r / geom.bw, pw = r - ph * geom.bw;
      valid = r < geom.rows && pw < geom.w_lim && ph < geom.h_lim;
      m = m_base + ph * geom.img_w + pw;
    } else {
      m = m_base + row0 + it * 4 + rs;
      valid = m < M;
    }
    if (valid) {
      int o = m;
      if (kOut16) {
        if (rpi > 0) o = (m / rpi) * static_cast<int>(p.out_item_stride) + (m % rpi);
        o += static_cast<int>(p.out_row_offset);
      }
      orow[it] = o;
      if (EPI == DWM_EPI_RESID) {
        item[it] = rpi > 0 ? m / rpi : 0;
        rrow[it] = p.resid_row_mod > 0 ? m % static_cast<int>(p.resid_row_mod)
                   : (p.resid_row_mod < 0 ? item[it] : m);
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
        // fused scatter to the peers' buffers over NVLink (same element offset)
        const long long eoff = static_cast<long long>(orow[it]) * p.ldo + ocol + c4 * 4;
        for (int q = 0; q < p.n_peers; ++q)
          *reinterpret_cast<uint2*>(reinterpret_cast<T*>(p.peer_out[q]) + eoff) = pk;
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
          float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
          if (p.gate)
            g = __ldg(reinterpret_cast<const float4*>(p.gate + static_cast<long long>(item[it]) * p.gate_ld + col));
          v.x = resid_elem(v.x, b.x, g.x, rq[it].x); v.y = resid_elem(v.y, b.y, g.y, rq[it].y);
          v.z = resid_elem(v.z, b.z, g.z, rq[it].z); v.w = resid_elem(v.w, b.w, g.w, rq[it].w);
          if (p.blend_x) {
            const float a = alpha[it], a1 = 1.0f - alpha[it];
            v.x = blend_elem(a, a1, bq[it].x, v.x); v.y = blend_elem(a, a1, bq[it].y, v.y);
            v.z = blend_elem(a, a1, bq[it].z, v.z); v.w = blend_elem(a, a1, bq[it].w, v.w);
          }
        }
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<long long>(orow[it]) * p.ldo + col) = v;
      }
    }
  };

  if constexpr (EPI == DWM_EPI_STORE || EPI == DWM_EPI_F32 || EPI == DWM_EPI_RESID) {
#pragma unroll 1
    const int ncols = geom.tile_cols > 0 ? geom.tile_cols : BN;
    for (int c = half; c * 32 < ncols; c += 2) {
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


// RESID epilogues read-modify-write a 128 KB fp32 tile whose HBM latency the 8 epilogue
// warps cannot cover with register prefetch alone.  While the MMAs of the tile are still
// running, each epilogue thread asks the L2 to fetch its row segment (512 B) of the
// residual (and blend) operand, so the later loads hit L2.
template <int EPI>
__device__ __forceinline__ void prefetch_resid_tile(const EpiParams& p, int m, int M, int n_tile0,
                                                    int N, int half, int half_cols = BN / 2) {
  if constexpr (EPI == DWM_EPI_RESID) {
    const int n0 = n_tile0 + half * half_cols;
    if (m < M && n0 < N) {
      const int cols = (N - n0) < half_cols ? (N - n0) : half_cols;
      const uint32_t bytes = static_cast<uint32_t>(cols) * 4u;
      if (p.resid) {
        if (p.resid_row_mod >= 0) {
          const long long rr = p.resid_row_mod > 0 ? m % static_cast<int>(p.resid_row_mod) : m;
          const float* src = p.resid + rr * p.ldr + n0;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
        }
      }
      if (p.blend_x) {
        const float* src = p.blend_x + static_cast<long long>(m) * p.ldx + n0;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
      }
    }
  }
}

// host side: fills EpiParams from the C-ABI struct
inline void fill_epi_params(EpiParams& p, const dwm_linear_args* a) {
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
  p.n_peers = a->n_peer_out;
  for (int i = 0; i < 8; ++i) p.peer_out[i] = i < a->n_peer_out ? a->peer_out[i] : nullptr;
}

}  // namespace dwm
