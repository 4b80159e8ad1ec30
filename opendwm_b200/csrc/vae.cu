// HBM-bound kernels of the CogVideoX temporal-VAE decoder (channels-last activations):
// GroupNorm statistics, the fused SpatialNorm3D (GroupNorm * conv_y(zq) + conv_b(zq)) +
// SiLU that writes the 16-bit, causally time-padded input of the next convolution, and
// the nearest-neighbour (space / space-time) upsampler.
#include <algorithm>

#include "common.cuh"
#include "../../include/dwm_b200.h"

namespace dwm {

// ---- GroupNorm statistics: sums[n][g] = (sum x, sum x^2) over (C/G channels, all pixels) ----
// block = 256 threads; thread handles float4 channel vector c4 = tid % (C/4) of pixels
// tid / (C/4), +stride...  Partial sums are combined per group in shared memory, then one
// double atomicAdd per (block, group).
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, long long pixels, int C, int G,
                                                       long long pixels_per_block, double* __restrict__ sums) {
  __shared__ float s_sum[64], s_sq[64];
  const int n = blockIdx.y;
  const int vec = C >> 2;
  const int tid = threadIdx.x;
  if (tid < 64) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
  __syncthreads();
  const int c4 = tid % vec;
  const int prow = tid / vec;
  const int rows_per_iter = 256 / vec;
  const long long p0 = static_cast<long long>(blockIdx.x) * pixels_per_block;
  long long p1 = p0 + pixels_per_block;
  if (p1 > pixels) p1 = pixels;
  float a = 0.f, b = 0.f;
  if (prow < rows_per_iter) {
    const float4* base = reinterpret_cast<const float4*>(x + static_cast<long long>(n) * pixels * C);
    for (long long p = p0 + prow; p < p1; p += rows_per_iter) {
      const float4 v = base[p * vec + c4];
      a += v.x + v.y + v.z + v.w;
      b += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  }
  const int g = (c4 * 4) / (C / G);
  atomicAdd(&s_sum[g], a);
  atomicAdd(&s_sq[g], b);
  __syncthreads();
  if (tid < G) {
    atomicAdd(&sums[(static_cast<long long>(n) * G + tid) * 2], static_cast<double>(s_sum[tid]));
    atomicAdd(&sums[(static_cast<long long>(n) * G + tid) * 2 + 1], static_cast<double>(s_sq[tid]));
  }
}

// wide variant: any C % 4 == 0 with C / 4 <= 1024 and any group size (UNet widths 320 ... 2560,
// 10 / 20 / 40 ... channels per group).  blockDim = (C/4) * k so that every thread keeps ONE
// channel quad for all its pixels: per-element register sums, then at most 4 shared atomics
// per thread and one double atomic per (block, group).
__global__ void __launch_bounds__(1024) gn_stats_wide_kernel(const float* __restrict__ x, long long pixels, int C, int G,
                                                             long long pixels_per_block, double* __restrict__ sums) {
  __shared__ float s_sum[64], s_sq[64];
  const int n = blockIdx.y, tid = threadIdx.x;
  const int vec = C >> 2;
  if (tid < 64) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
  __syncthreads();
  const int c4 = tid % vec, prow = tid / vec, k = blockDim.x / vec;
  const long long p0 = static_cast<long long>(blockIdx.x) * pixels_per_block;
  long long p1 = p0 + pixels_per_block;
  if (p1 > pixels) p1 = pixels;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
  const float4* base = reinterpret_cast<const float4*>(x + static_cast<long long>(n) * pixels * C);
  for (long long p = p0 + prow; p < p1; p += k) {
    const float4 v = base[p * vec + c4];
    s0 += v.x; q0 += v.x * v.x;
    s1 += v.y; q1 += v.y * v.y;
    s2 += v.z; q2 += v.z * v.z;
    s3 += v.w; q3 += v.w * v.w;
  }
  const int cg = C / G;
  const float ss[4] = {s0, s1, s2, s3}, qq[4] = {q0, q1, q2, q3};
  int cur = (c4 * 4) / cg;
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int g = (c4 * 4 + e) / cg;
    if (g != cur) {
      atomicAdd(&s_sum[cur], a); atomicAdd(&s_sq[cur], b);
      cur = g; a = 0.f; b = 0.f;
    }
    a += ss[e]; b += qq[e];
  }
  atomicAdd(&s_sum[cur], a); atomicAdd(&s_sq[cur], b);
  __syncthreads();
  if (tid < G) {
    atomicAdd(&sums[(static_cast<long long>(n) * G + tid) * 2], static_cast<double>(s_sum[tid]));
    atomicAdd(&sums[(static_cast<long long>(n) * G + tid) * 2 + 1], static_cast<double>(s_sq[tid]));
  }
}

// last-resort variant (C / 4 > 1024): any C, any group size (scalar loads; UNet widths 320 / 960 / 1920 ...)
__global__ void __launch_bounds__(256) gn_stats_generic_kernel(const float* __restrict__ x, long long pixels, int C,
                                                               int G, long long pixels_per_block,
                                                               double* __restrict__ sums) {
  __shared__ float s_sum[64], s_sq[64];
  const int n = blockIdx.y, tid = threadIdx.x;
  if (tid < 64) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
  __syncthreads();
  const int cg = C / G;
  const long long e0 = static_cast<long long>(blockIdx.x) * pixels_per_block * C;
  long long e1 = e0 + pixels_per_block * C;
  if (e1 > pixels * C) e1 = pixels * C;
  const float* base = x + static_cast<long long>(n) * pixels * C;
  // consecutive threads read consecutive elements; accumulate runs of equal group locally
  int cur_g = -1;
  float a = 0.f, b = 0.f;
  for (long long e = e0 + tid; e < e1; e += 256) {
    const int g = static_cast<int>(e % C) / cg;
    if (g != cur_g) {
      if (cur_g >= 0) { atomicAdd(&s_sum[cur_g], a); atomicAdd(&s_sq[cur_g], b); }
      cur_g = g; a = 0.f; b = 0.f;
    }
    const float v = base[e];
    a += v; b += v * v;
  }
  if (cur_g >= 0) { atomicAdd(&s_sum[cur_g], a); atomicAdd(&s_sq[cur_g], b); }
  __syncthreads();
  if (tid < G) {
    atomicAdd(&sums[(static_cast<long long>(n) * G + tid) * 2], static_cast<double>(s_sum[tid]));
    atomicAdd(&sums[(static_cast<long long>(n) * G + tid) * 2 + 1], static_cast<double>(s_sq[tid]));
  }
}

// ---- SpatialNorm3D / GroupNorm apply (+SiLU) -> 16-bit, written at a frame offset ----
struct SnParams {
  const float* x; int nb, T, H, W, C, G;
  const double* sums; float eps;
  const float* gamma; const float* beta;
  const float* zy; const float* zb; int Tz, hz, wz;
  int silu;
  void* out; int out_T, out_t0;
};

// grid = (chunks, nb): a block works on one chunk (16 float4 per thread) of ONE image, so the (mean, rstd) of its
// G groups are finalised once per block from the fp64 sums (first G threads, shared memory)
// instead of per thread — the fp64 divisions made the per-thread version XU-pipe bound.

// Per-thread state is set up ONCE: with blockDim a multiple of C/4 a thread keeps the same
// channel quad for all its elements, so gamma / beta / (mean, rstd) are registers, the pixel
// position (t, h, w) advances by a constant number of pixels per iteration and is tracked with
// carries instead of 64-bit divisions, the latent-grid coordinates are shifts when the sizes
// are powers of two apart (always in the VAEs) and the frame map is a 64-entry shared table.
// The r01 kernel spent ~10 integer divisions per float4 and ran at 18 % of DRAM peak
// (profiles/r01_ncu_kernels_summary.txt); this one is a plain stream.
template <typename T>
__global__ void __launch_bounds__(1024) spatialnorm_kernel(const SnParams p, const int chunk) {
  // blockDim.x is a multiple of C/4 whenever C/4 <= 1024 (host side), `chunk` = blockDim.x * 16
  __shared__ float2 s_stat[64];
  __shared__ int s_tz[64];
  const int NT = static_cast<int>(blockDim.x);
  const int vec = p.C >> 2;
  const int n = blockIdx.y;
  const int cg = p.C / p.G;
  const long long per_img = static_cast<long long>(p.T) * p.H * p.W * vec;
  if (threadIdx.x < p.G) {
    const double cnt = static_cast<double>(cg) * p.T * p.H * p.W;
    const double s = p.sums[(static_cast<long long>(n) * p.G + threadIdx.x) * 2];
    const double ss = p.sums[(static_cast<long long>(n) * p.G + threadIdx.x) * 2 + 1];
    const double mean_d = s / cnt;
    s_stat[threadIdx.x] = make_float2(static_cast<float>(mean_d),
                                      rsqrtf(static_cast<float>(ss / cnt - mean_d * mean_d) + p.eps));
  }
  if (p.zy && threadIdx.x < 64 && threadIdx.x < p.T) {
    // nearest-neighbour frame of the latent grid; odd T > 1 treats the first frame apart
    const int t = threadIdx.x;
    s_tz[t] = (p.T > 1 && (p.T & 1)) ? (t == 0 ? 0 : 1 + ((t - 1) * (p.Tz - 1)) / (p.T - 1))
                                     : (t * p.Tz) / p.T;
  }
  __syncthreads();
  const long long i0 = static_cast<long long>(blockIdx.x) * chunk;
  long long i1 = i0 + chunk;
  if (i1 > per_img) i1 = per_img;
  const float4* xin = reinterpret_cast<const float4*>(p.x) + static_cast<long long>(n) * per_img;
  const float4* g4 = reinterpret_cast<const float4*>(p.gamma);
  const float4* b4 = reinterpret_cast<const float4*>(p.beta);
  const bool fast = (NT % vec) == 0 && (chunk % vec) == 0 && p.T <= 64;
  if (fast) {
    const int c4 = threadIdx.x % vec;
    const int pstep = NT / vec;                                   // pixels per iteration
    long long pix = i0 / vec + threadIdx.x / vec;                 // i0 % vec == 0
    int w = static_cast<int>(pix % p.W);
    long long r = pix / p.W;
    int h = static_cast<int>(r % p.H);
    int t = static_cast<int>(r / p.H);
    const float4 ga = __ldg(g4 + c4), be = __ldg(b4 + c4);
    // the four channels of the quad may sit in different groups (UNet: 10 / 20 / 40 channels
    // per group): one (mean, rstd) per element, all in registers
    const float2 st0 = s_stat[(c4 * 4) / cg], st1 = s_stat[(c4 * 4 + 1) / cg];
    const float2 st2 = s_stat[(c4 * 4 + 2) / cg], st3 = s_stat[(c4 * 4 + 3) / cg];
    // (x - mean) * (rstd * gamma) + beta: the subtraction stays first (no cancellation in a
    // pre-folded offset when |mean| >> std)
    const float4 aa = make_float4(st0.y * ga.x, st1.y * ga.y, st2.y * ga.z, st3.y * ga.w);
    const float4 mu = make_float4(st0.x, st1.x, st2.x, st3.x);
    // latent-grid coordinates: shifts when H = hz << k (else a division per element)
    int hs = -1, ws = -1;
    if (p.zy) {
      for (int k = 0; k < 8; ++k) {
        if ((p.hz << k) == p.H) hs = k;
        if ((p.wz << k) == p.W) ws = k;
      }
    }
    const long long zn = static_cast<long long>(n) * p.Tz;
    const long long on = static_cast<long long>(n) * p.out_T + p.out_t0;
    for (long long i = i0 + threadIdx.x; i < i1; i += NT) {
      float4 v = xin[i];
      v.x = fmaf(v.x - mu.x, aa.x, be.x); v.y = fmaf(v.y - mu.y, aa.y, be.y);
      v.z = fmaf(v.z - mu.z, aa.z, be.z); v.w = fmaf(v.w - mu.w, aa.w, be.w);
      if (p.zy) {
        const int hq = hs >= 0 ? (h >> hs) : (h * p.hz) / p.H;
        const int wq = ws >= 0 ? (w >> ws) : (w * p.wz) / p.W;
        const long long zi = (((zn + s_tz[t]) * p.hz + hq) * p.wz + wq) * vec + c4;
        const float4 y = __ldg(reinterpret_cast<const float4*>(p.zy) + zi);
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.zb) + zi);
        v.x = fmaf(v.x, y.x, b.x); v.y = fmaf(v.y, y.y, b.y);
        v.z = fmaf(v.z, y.z, b.z); v.w = fmaf(v.w, y.w, b.w);
      }
      if (p.silu) { v.x = silu(v.x); v.y = silu(v.y); v.z = silu(v.z); v.w = silu(v.w); }
      const long long o = (((on + t) * p.H + h) * p.W + w) * vec + c4;
      uint2 pk; pk.x = Cvt<T>::pack2(v.x, v.y); pk.y = Cvt<T>::pack2(v.z, v.w);
      reinterpret_cast<uint2*>(p.out)[o] = pk;
      w += pstep;
      while (w >= p.W) { w -= p.W; if (++h == p.H) { h = 0; ++t; } }
    }
    return;
  }
  // general path (C / 4 > 1024 or more than 64 frames)
  for (long long i = i0 + threadIdx.x; i < i1; i += NT) {
    const int c4 = static_cast<int>(i % vec);
    long long r = i / vec;
    const int w = static_cast<int>(r % p.W); r /= p.W;
    const int h = static_cast<int>(r % p.H);
    const int t = static_cast<int>(r / p.H);
    float4 v = xin[i];
    const float4 ga = __ldg(g4 + c4);
    const float4 be = __ldg(b4 + c4);
    if ((cg & 3) == 0) {
      const float2 st = s_stat[(c4 * 4) / cg];
      v.x = (v.x - st.x) * st.y * ga.x + be.x;
      v.y = (v.y - st.x) * st.y * ga.y + be.y;
      v.z = (v.z - st.x) * st.y * ga.z + be.z;
      v.w = (v.w - st.x) * st.y * ga.w + be.w;
    } else {   // a float4 may straddle groups (e.g. 10 or 2 channels per group)
      float* ve = reinterpret_cast<float*>(&v);
      const float* gae = reinterpret_cast<const float*>(&ga);
      const float* bee = reinterpret_cast<const float*>(&be);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 st = s_stat[(c4 * 4 + e) / cg];
        ve[e] = (ve[e] - st.x) * st.y * gae[e] + bee[e];
      }
    }
    if (p.zy) {
      int tz;
      if (p.T > 1 && (p.T & 1)) tz = t == 0 ? 0 : 1 + ((t - 1) * (p.Tz - 1)) / (p.T - 1);
      else tz = (t * p.Tz) / p.T;
      const int hq = (h * p.hz) / p.H, wq = (w * p.wz) / p.W;
      const long long zi = (((static_cast<long long>(n) * p.Tz + tz) * p.hz + hq) * p.wz + wq) * vec + c4;
      const float4 y = __ldg(reinterpret_cast<const float4*>(p.zy) + zi);
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.zb) + zi);
      v.x = v.x * y.x + b.x; v.y = v.y * y.y + b.y; v.z = v.z * y.z + b.z; v.w = v.w * y.w + b.w;
    }
    if (p.silu) { v.x = silu(v.x); v.y = silu(v.y); v.z = silu(v.z); v.w = silu(v.w); }
    const long long o = (((static_cast<long long>(n) * p.out_T + p.out_t0 + t) * p.H + h) * p.W + w) * vec + c4;
    uint2 pk; pk.x = Cvt<T>::pack2(v.x, v.y); pk.y = Cvt<T>::pack2(v.z, v.w);
    reinterpret_cast<uint2*>(p.out)[o] = pk;
  }
}

// ---- nearest upsample x2 in space, optionally in time (CogVideoXUpsample3D rules) ----
template <typename T>
__global__ void __launch_bounds__(256) upsample_kernel(const float* __restrict__ x, int nb, int Ti, int H, int W, int C,
                                                       int To, int mode, T* __restrict__ out) {
  // mode 0: space only; 1: space+time all frames; 2: first frame space only, rest space+time
  const int vec = C >> 2;
  const long long total = static_cast<long long>(nb) * To * (2 * H) * (2 * W) * vec;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = static_cast<int>(i % vec);
  long long r = i / vec;
  const int w = static_cast<int>(r % (2 * W)); r /= 2 * W;
  const int h = static_cast<int>(r % (2 * H)); r /= 2 * H;
  const int t = static_cast<int>(r % To);
  const int n = static_cast<int>(r / To);
  int ti;
  if (mode == 0) ti = t;
  else if (mode == 1) ti = t >> 1;
  else ti = t == 0 ? 0 : 1 + ((t - 1) >> 1);
  const float4 v = reinterpret_cast<const float4*>(x)[(((static_cast<long long>(n) * Ti + ti) * H + (h >> 1)) * W + (w >> 1)) * vec + c4];
  uint2 pk; pk.x = Cvt<T>::pack2(v.x, v.y); pk.y = Cvt<T>::pack2(v.z, v.w);
  reinterpret_cast<uint2*>(out)[i] = pk;
}

}  // namespace dwm

using namespace dwm;

extern "C" int dwm_b200_groupnorm_stats(const float* x, int64_t nb, int64_t pixels, int C, int groups,
                                        double* sums, dwm_stream_t stream) {
  DWM_REQUIRE(x && sums && nb > 0 && pixels > 0, "dwm_b200_groupnorm_stats: bad arguments");
  DWM_REQUIRE(C % 4 == 0 && groups > 0 && groups <= 64 && C % groups == 0,
              "dwm_b200_groupnorm_stats: need C %% 4 == 0, C %% groups == 0, groups <= 64 (got C=%d, groups=%d)", C, groups);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DWM_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * nb * groups, s));
  const int vec = C / 4;
  const bool fast = (C / groups) % 4 == 0 && vec <= 256 && 256 % vec == 0;
  // pixels per block: about 4 blocks per SM over the whole launch, whole thread-rows per block
  auto pick = [&](long long rows_per_iter) {
    const long long target = std::max<long long>(1, 592 / nb);
    const long long chunks = std::min((pixels + rows_per_iter - 1) / rows_per_iter, target);
    const long long ppb = (pixels + chunks - 1) / chunks;
    return (ppb + rows_per_iter - 1) / rows_per_iter * rows_per_iter;
  };
  if (fast) {
    const long long ppb = pick(256 / vec);
    dim3 grid(static_cast<unsigned>((pixels + ppb - 1) / ppb), static_cast<unsigned>(nb));
    gn_stats_kernel<<<grid, 256, 0, s>>>(x, pixels, C, groups, ppb, sums);
  } else if (vec <= 1024) {
    const int k = std::max(1, 256 / vec);
    const long long ppb = pick(k);
    dim3 grid(static_cast<unsigned>((pixels + ppb - 1) / ppb), static_cast<unsigned>(nb));
    gn_stats_wide_kernel<<<grid, vec * k, 0, s>>>(x, pixels, C, groups, ppb, sums);
  } else {
    const long long ppb = 64;
    dim3 grid(static_cast<unsigned>((pixels + ppb - 1) / ppb), static_cast<unsigned>(nb));
    gn_stats_generic_kernel<<<grid, 256, 0, s>>>(x, pixels, C, groups, ppb, sums);
  }
  DWM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dwm_b200_spatialnorm_silu(const float* x, int64_t nb, int64_t T, int64_t H, int64_t W, int C,
                                         int groups, const double* sums, float eps, const float* gamma,
                                         const float* beta, const float* zy, const float* zb, int Tz, int hz,
                                         int wz, int apply_silu, void* out, int64_t out_T, int64_t out_t0,
                                         int dtype, dwm_stream_t stream) {
  DWM_REQUIRE(x && sums && gamma && beta && out, "dwm_b200_spatialnorm_silu: null pointer");
  DWM_REQUIRE(C % 4 == 0 && groups > 0 && C % groups == 0, "dwm_b200_spatialnorm_silu: bad C/groups");
  DWM_REQUIRE((zy == nullptr) == (zb == nullptr), "dwm_b200_spatialnorm_silu: zy and zb go together");
  DWM_REQUIRE(out_t0 >= 0 && out_t0 + T <= out_T, "dwm_b200_spatialnorm_silu: frame window outside out buffer");
  SnParams p;
  p.x = x; p.nb = (int)nb; p.T = (int)T; p.H = (int)H; p.W = (int)W; p.C = C; p.G = groups;
  p.sums = sums; p.eps = eps; p.gamma = gamma; p.beta = beta; p.zy = zy; p.zb = zb;
  p.Tz = Tz; p.hz = hz; p.wz = wz; p.silu = apply_silu; p.out = out; p.out_T = (int)out_T; p.out_t0 = (int)out_t0;
  DWM_REQUIRE(groups <= 64 && nb <= 65535, "dwm_b200_spatialnorm_silu: groups <= 64 and nb <= 65535 required");
  const long long per_img = T * H * W * (C / 4);
  // block = the largest multiple of C/4 that fits 256 threads (or C/4 itself up to 1024), so a
  // thread keeps one channel quad; 16 float4 per thread
  const int vec = C / 4;
  int threads = 256;
  if (vec <= 1024) threads = vec <= 256 ? (256 / vec) * vec : vec;
  const int chunk = threads * 16;
  dim3 grid(static_cast<unsigned>((per_img + chunk - 1) / chunk), static_cast<unsigned>(nb));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == DWM_BF16) spatialnorm_kernel<__nv_bfloat16><<<grid, threads, 0, s>>>(p, chunk);
  else if (dtype == DWM_F16) spatialnorm_kernel<__half><<<grid, threads, 0, s>>>(p, chunk);
  else { set_last_error("dwm_b200_spatialnorm_silu: bad dtype"); return -1; }
  DWM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dwm_b200_upsample_nearest(const float* x, int64_t nb, int64_t T, int64_t H, int64_t W, int C,
                                         int compress_time, void* out, int dtype, dwm_stream_t stream) {
  DWM_REQUIRE(x && out && C % 4 == 0, "dwm_b200_upsample_nearest: bad arguments");
  int mode = 0;
  long long To = T;
  if (compress_time && T > 1) {
    if (T % 2 == 1) { mode = 2; To = 1 + 2 * (T - 1); }
    else { mode = 1; To = 2 * T; }
  }
  const long long total = nb * To * 2 * H * 2 * W * (C / 4);
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == DWM_BF16)
    upsample_kernel<<<grid, 256, 0, s>>>(x, (int)nb, (int)T, (int)H, (int)W, C, (int)To, mode, reinterpret_cast<__nv_bfloat16*>(out));
  else if (dtype == DWM_F16)
    upsample_kernel<<<grid, 256, 0, s>>>(x, (int)nb, (int)T, (int)H, (int)W, C, (int)To, mode, reinterpret_cast<__half*>(out));
  else { set_last_error("dwm_b200_upsample_nearest: bad dtype"); return -1; }
  DWM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
