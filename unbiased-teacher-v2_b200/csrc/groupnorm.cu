// GroupNorm(32, 256) + ReLU for the FCOS towers, NHWC bf16 (reference: nn.GroupNorm + nn.ReLU,
// ubteacher/modeling/fcos/fcos.py:263-264,283). Eight channels per group == one 16-byte vector per
// (pixel, group), so every kernel is a flat coalesced uint4 stream; statistics are accumulated in
// fp64 atomics per (image, group).
//
// Algorithmic bytes per element (bf16): fwd = read x twice + write y = 6 B; bwd = read x, dy twice +
// write dx = 10 B.
#include "ut2_internal.h"
#include "sm100_ptx.cuh"
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

namespace {
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void unpack8(const uint4 v, float (&f)[8]) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(u[j] << 16);
    f[2 * j + 1] = __uint_as_float(u[j] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    o[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

constexpr int GN_G = 32;          // groups (== vectors per pixel)
constexpr int GN_ROWS = 8;        // pixel rows per block iteration (256 threads)
#ifndef GN_BWD_OCC
#define GN_BWD_OCC 2
#endif
constexpr int GN_UNROLL = 4;      // independent 16-byte loads in flight per thread

// One launch covers up to 5 pyramid levels of a level-major buffer (level l = [N, HW_l, 256] starting at pixel row
// row_off[l]); a plain [N, HW, 256] tensor is the 1-level case. Blocks are enumerated (level, image, pixel chunk);
// statistics live at stats[(l * N + n) * 32 + g].
constexpr int GN_MAXL = 5;
struct GnLevels {
  int num, N;
  int HW[GN_MAXL], row_off[GN_MAXL], ppb[GN_MAXL], bpi[GN_MAXL];
  int blk_off[GN_MAXL + 1];
};
struct GnBlock {
  int n_stat;        // l * N + n
  int HW, p0, p1;
  size_t base;       // first uint4 of image n in level l
};
__device__ __forceinline__ GnBlock gn_locate(const GnLevels& lv) {
  const int b = blockIdx.x;
  int l = 0;
#pragma unroll
  for (int i = 1; i < GN_MAXL; ++i)
    if (i < lv.num && b >= lv.blk_off[i]) l = i;
  const int rem = b - lv.blk_off[l];
  const int n = rem / lv.bpi[l], chunk = rem - n * lv.bpi[l];
  GnBlock r;
  r.n_stat = l * lv.N + n;
  r.HW = lv.HW[l];
  r.p0 = chunk * lv.ppb[l];
  r.p1 = min(r.HW, r.p0 + lv.ppb[l]);
  r.base = ((size_t)lv.row_off[l] + (size_t)n * r.HW) * GN_G;
  return r;
}

// stats[n][g] = {sum, sumsq} over HW x 8 channels
__global__ void __launch_bounds__(256)
gn_stats_kernel(const bf16* __restrict__ x, double* __restrict__ stats, GnLevels lv) {
  ut2::griddep_wait();
  ut2::griddep_launch();
  const GnBlock B = gn_locate(lv);
  const int n = B.n_stat, g = threadIdx.x & 31, row = threadIdx.x >> 5;
  const int p0 = B.p0, p1 = B.p1;
  const uint4* xv = reinterpret_cast<const uint4*>(x) + B.base;
  float s1 = 0.f, s2 = 0.f;
  for (int p = p0 + row; p < p1; p += GN_ROWS * GN_UNROLL) {
    uint4 v[GN_UNROLL];
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {       // all loads first: GN_UNROLL x 16 B in flight per thread
      const int pp = p + u * GN_ROWS;
      v[u] = pp < p1 ? __ldg(xv + (size_t)pp * GN_G + g) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1 += f[j]; s2 = fmaf(f[j], f[j], s2); }
    }
  }
  __shared__ float red[GN_ROWS][GN_G][2];
  red[row][g][0] = s1;
  red[row][g][1] = s2;
  __syncthreads();
  if (row == 0) {
    double d1 = 0.0, d2 = 0.0;
    for (int k = 0; k < GN_ROWS; ++k) { d1 += red[k][g][0]; d2 += red[k][g][1]; }
    atomicAdd(stats + ((size_t)n * GN_G + g) * 2, d1);
    atomicAdd(stats + ((size_t)n * GN_G + g) * 2 + 1, d2);
  }
}

__device__ __forceinline__ void mean_rstd(const double* stats, int n, int g, int HW, float eps, float& mean,
                                          float& rstd) {
  const double m = (double)HW * 8.0;
  const double mu = stats[((size_t)n * GN_G + g) * 2] / m;
  double var = stats[((size_t)n * GN_G + g) * 2 + 1] / m - mu * mu;
  if (var < 0.0) var = 0.0;
  mean = (float)mu;
  rstd = (float)(1.0 / sqrt(var + (double)eps));
}

__global__ void __launch_bounds__(256)
gn_apply_kernel(const bf16* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, bf16* __restrict__ y, GnLevels lv, int relu) {
  ut2::griddep_wait();
  ut2::griddep_launch();
  const GnBlock B = gn_locate(lv);
  const int n = B.n_stat, HW = B.HW, g = threadIdx.x & 31, row = threadIdx.x >> 5;
  float mean, rstd;
  mean_rstd(stats, n, g, HW, eps, mean, rstd);
  float ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ga[j] = gamma[g * 8 + j] * rstd; be[j] = beta[g * 8 + j] - mean * ga[j]; }
  const int p0 = B.p0, p1 = B.p1;
  const uint4* xv = reinterpret_cast<const uint4*>(x) + B.base;
  uint4* yv = reinterpret_cast<uint4*>(y) + B.base;
  for (int p = p0 + row; p < p1; p += GN_ROWS * GN_UNROLL) {
    uint4 v[GN_UNROLL];
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      const int pp = p + u * GN_ROWS;
      v[u] = pp < p1 ? __ldg(xv + (size_t)pp * GN_G + g) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      const int pp = p + u * GN_ROWS;
      if (pp < p1) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[j] = fmaf(f[j], ga[j], be[j]);
          if (relu) f[j] = fmaxf(f[j], 0.f);
        }
        yv[(size_t)pp * GN_G + g] = pack8(f);
      }
    }
  }
}

// pass 1 of backward: per-(n, group) sums s1 = sum g*gamma, s2 = sum g*gamma*xhat; per-channel dgamma/dbeta
__global__ void __launch_bounds__(256, GN_BWD_OCC)
gn_bwd_reduce_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const double* __restrict__ stats,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                     double* __restrict__ ws, float* __restrict__ dgamma, float* __restrict__ dbeta, GnLevels lv,
                     int relu) {
  ut2::griddep_wait();
  ut2::griddep_launch();
  const GnBlock B = gn_locate(lv);
  const int n = B.n_stat, HW = B.HW, g = threadIdx.x & 31, row = threadIdx.x >> 5;
  float mean, rstd;
  mean_rstd(stats, n, g, HW, eps, mean, rstd);
  float ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ga[j] = gamma[g * 8 + j]; be[j] = beta[g * 8 + j]; }
  const int p0 = B.p0, p1 = B.p1;
  const uint4* xv = reinterpret_cast<const uint4*>(x) + B.base;
  const uint4* dv = reinterpret_cast<const uint4*>(dy) + B.base;
  float dg[8], db[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { dg[j] = 0.f; db[j] = 0.f; }
  // Two register batches in ping-pong: the loads of the next GN_UNROLL rows are in flight while this batch is reduced (ncu on the
  // single-buffered loop: 24 % of the warp slots active and ~60 % of the samples waiting on these loads).
  auto load = [&](uint4 (&vx)[GN_UNROLL], uint4 (&vd)[GN_UNROLL], int p) {
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      const int pp = p + u * GN_ROWS;
      vx[u] = pp < p1 ? __ldg(xv + (size_t)pp * GN_G + g) : make_uint4(0, 0, 0, 0);
      vd[u] = pp < p1 ? __ldg(dv + (size_t)pp * GN_G + g) : make_uint4(0, 0, 0, 0);   // zero gradient: no contribution
    }
  };
  auto reduce = [&](const uint4 (&vx)[GN_UNROLL], const uint4 (&vd)[GN_UNROLL]) {
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      float f[8], d[8];
      unpack8(vx[u], f);
      unpack8(vd[u], d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (f[j] - mean) * rstd;
        float gj = d[j];
        if (relu && !(fmaf(xh, ga[j], be[j]) > 0.f)) gj = 0.f;
        dg[j] = fmaf(gj, xh, dg[j]);
        db[j] += gj;
        const float gg = gj * ga[j];
        s1 += gg;
        s2 = fmaf(gg, xh, s2);
      }
    }
  };
  {
    constexpr int STEP = GN_ROWS * GN_UNROLL;
    uint4 ax[GN_UNROLL], ad[GN_UNROLL], bx[GN_UNROLL], bd[GN_UNROLL];
    int p = p0 + row;
    if (p < p1) {
      load(ax, ad, p);
#pragma unroll 1
      while (true) {
        load(bx, bd, p + STEP);
        reduce(ax, ad);
        p += STEP;
        if (p >= p1) break;
        load(ax, ad, p + STEP);
        reduce(bx, bd);
        p += STEP;
        if (p >= p1) break;
      }
    }
  }
  __shared__ float red[GN_ROWS][GN_G][18];
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[row][g][j] = dg[j]; red[row][g][8 + j] = db[j]; }
  red[row][g][16] = s1;
  red[row][g][17] = s2;
  __syncthreads();
  if (row == 0) {
    float acc[18];
#pragma unroll
    for (int j = 0; j < 18; ++j) acc[j] = 0.f;
    for (int k = 0; k < GN_ROWS; ++k)
#pragma unroll
      for (int j = 0; j < 18; ++j) acc[j] += red[k][g][j];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(dgamma + g * 8 + j, acc[j]);
      atomicAdd(dbeta + g * 8 + j, acc[8 + j]);
    }
    atomicAdd(ws + ((size_t)n * GN_G + g) * 2, (double)acc[16]);
    atomicAdd(ws + ((size_t)n * GN_G + g) * 2 + 1, (double)acc[17]);
  }
}

// pass 2: dx = rstd * (g*gamma - s1/m - xhat*s2/m)
__global__ void __launch_bounds__(256, GN_BWD_OCC)
gn_bwd_apply_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const double* __restrict__ stats,
                    const double* __restrict__ ws, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float eps, bf16* __restrict__ dx, float* __restrict__ dbias, GnLevels lv, int relu) {
  ut2::griddep_wait();
  ut2::griddep_launch();
  const GnBlock B = gn_locate(lv);
  const int n = B.n_stat, HW = B.HW, g = threadIdx.x & 31, row = threadIdx.x >> 5;
  float mean, rstd;
  mean_rstd(stats, n, g, HW, eps, mean, rstd);
  const double m = (double)HW * 8.0;
  const float a1 = (float)(ws[((size_t)n * GN_G + g) * 2] / m);
  const float a2 = (float)(ws[((size_t)n * GN_G + g) * 2 + 1] / m);
  float ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ga[j] = gamma[g * 8 + j]; be[j] = beta[g * 8 + j]; }
  const int p0 = B.p0, p1 = B.p1;
  const uint4* xv = reinterpret_cast<const uint4*>(x) + B.base;
  const uint4* dv = reinterpret_cast<const uint4*>(dy) + B.base;
  uint4* ov = reinterpret_cast<uint4*>(dx) + B.base;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // column sums of dx (the bias gradient of the conv before)
  auto load = [&](uint4 (&vx)[GN_UNROLL], uint4 (&vd)[GN_UNROLL], int p) {
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      const int pp = p + u * GN_ROWS;
      vx[u] = pp < p1 ? __ldg(xv + (size_t)pp * GN_G + g) : make_uint4(0, 0, 0, 0);
      vd[u] = pp < p1 ? __ldg(dv + (size_t)pp * GN_G + g) : make_uint4(0, 0, 0, 0);
    }
  };
  auto apply = [&](const uint4 (&vx)[GN_UNROLL], const uint4 (&vd)[GN_UNROLL], int p) {
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      const int pp = p + u * GN_ROWS;
      if (pp < p1) {
        float f[8], d[8];
        unpack8(vx[u], f);
        unpack8(vd[u], d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (f[j] - mean) * rstd;
          float gj = d[j];
          if (relu && !(fmaf(xh, ga[j], be[j]) > 0.f)) gj = 0.f;
          f[j] = rstd * (gj * ga[j] - a1 - xh * a2);
        }
        const uint4 o = pack8(f);
        ov[(size_t)pp * GN_G + g] = o;
        if (dbias) {                 // sum what was actually stored (bf16), like a separate column-sum pass would
          float r8[8];
          unpack8(o, r8);
#pragma unroll
          for (int j = 0; j < 8; ++j) cs[j] += r8[j];
        }
      }
    }
  };
  {
    constexpr int STEP = GN_ROWS * GN_UNROLL;      // ping-pong register batches, as in the reduce pass
    uint4 ax[GN_UNROLL], ad[GN_UNROLL], bx[GN_UNROLL], bd[GN_UNROLL];
    int p = p0 + row;
    if (p < p1) {
      load(ax, ad, p);
#pragma unroll 1
      while (true) {
        load(bx, bd, p + STEP);
        apply(ax, ad, p);
        p += STEP;
        if (p >= p1) break;
        load(ax, ad, p + STEP);
        apply(bx, bd, p);
        p += STEP;
        if (p >= p1) break;
      }
    }
  }
  if (dbias) {
    __shared__ float red[GN_ROWS][GN_G][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) red[row][g][j] = cs[j];
    __syncthreads();
    if (row == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = 0.f;
        for (int k = 0; k < GN_ROWS; ++k) t += red[k][g][j];
        atomicAdd(dbias + g * 8 + j, t);
      }
    }
  }
}

static int gn_min_px() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("UT2_GN_MINPX"); v = e ? atoi(e) : 256; if (v < 8) v = 8; }
  return v;
}

inline int fill_gn_levels(GnLevels& lv, int num_levels, const int* hws, int N, bool bwd = false) {
  if (num_levels < 1 || num_levels > GN_MAXL) return -1;
  lv.num = num_levels;
  lv.N = N;
  long long total = 0;
  for (int l = 0; l < num_levels; ++l) total += hws[l];
  // aim for ~148*4 blocks in total, spread over the levels by size, at least 256 pixels each (UT2_GN_MINPX): every block ends
  // in ~500 fp32 / fp64 atomics on the same few addresses, which dominates small batches when the blocks are short
  static int mult_f = -1, mult_b = -1;
  if (mult_f < 0) {
    const char* e = getenv("UT2_GN_BLOCKS"); mult_f = e ? atoi(e) : 4; if (mult_f < 1) mult_f = 1;
    e = getenv("UT2_GN_BLOCKS_BWD"); mult_b = e ? atoi(e) : 2; if (mult_b < 1) mult_b = 1;
  }
  const int mult = bwd ? mult_b : mult_f;
  const long long target = (total * N + 148 * mult - 1) / (148 * mult);
  int row = 0;
  lv.blk_off[0] = 0;
  for (int l = 0; l < GN_MAXL; ++l) {
    if (l < num_levels) {
      const int HW = hws[l];
      int ppb = (int)(target < gn_min_px() ? gn_min_px() : target);
      ppb = (ppb + GN_ROWS - 1) / GN_ROWS * GN_ROWS;
      lv.HW[l] = HW; lv.row_off[l] = row; lv.ppb[l] = ppb; lv.bpi[l] = (HW + ppb - 1) / ppb;
      lv.blk_off[l + 1] = lv.blk_off[l] + lv.bpi[l] * N;
      row += N * HW;
    } else {
      lv.HW[l] = 1; lv.row_off[l] = 0; lv.ppb[l] = GN_ROWS; lv.bpi[l] = 1;
      lv.blk_off[l + 1] = lv.blk_off[l];
    }
  }
  return 0;
}
}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

static double gn_est_us(int num_levels, const int* hws, int N, double bytes_per_elem) {
  double px = 0;
  for (int l = 0; l < num_levels; ++l) px += hws[l];
  return ut2_est_us(0.0, px * N * 256.0 * bytes_per_elem);
}

static int gn_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y, double* stats, int num_levels,
                  const int* hws, int N, int C, int G, int relu, void* stream) {
  if (C != 256 || G != 32) return ut2_fail(-2, "groupnorm: only GroupNorm(32, 256) is supported");
  GnLevels lv;
  if (fill_gn_levels(lv, num_levels, hws, N)) return ut2_fail(-3, "groupnorm: 1..5 levels");
  cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * num_levels * N * GN_G * 2, STREAM);
  if (e != cudaSuccess) return ut2_fail((int)e, "groupnorm: memset failed");
  const int grid = lv.blk_off[num_levels];
  gn_stats_kernel<<<grid, 256, 0, STREAM>>>(static_cast<const bf16*>(x), stats, lv);      // follows a memset node: plain launch
  ut2_launch_pdl(gn_apply_kernel, dim3(grid), dim3(256), 0, STREAM, gn_est_us(num_levels, hws, N, 4.0), static_cast<const bf16*>(x), (const double*)stats, gamma, beta, eps,
                 static_cast<bf16*>(y), lv, relu);
  return ut2_check_launch("groupnorm_fwd");
}

static int gn_bwd(const void* dy, const void* x, const double* stats, const float* gamma, const float* beta, float eps,
                  void* dx, float* dgamma, float* dbeta, float* dbias_prev, double* ws, int num_levels, const int* hws, int N,
                  int C, int G, int relu, void* stream) {
  if (C != 256 || G != 32) return ut2_fail(-2, "groupnorm: only GroupNorm(32, 256) is supported");
  GnLevels lv;
  if (fill_gn_levels(lv, num_levels, hws, N, true)) return ut2_fail(-3, "groupnorm: 1..5 levels");
  cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(double) * num_levels * N * GN_G * 2, STREAM);
  if (e != cudaSuccess) return ut2_fail((int)e, "groupnorm: memset failed");
  const int grid = lv.blk_off[num_levels];
  gn_bwd_reduce_kernel<<<grid, 256, 0, STREAM>>>(static_cast<const bf16*>(dy), static_cast<const bf16*>(x), stats, gamma, beta,
                                                 eps, ws, dgamma, dbeta, lv, relu);
  ut2_launch_pdl(gn_bwd_apply_kernel, dim3(grid), dim3(256), 0, STREAM, gn_est_us(num_levels, hws, N, 6.0), static_cast<const bf16*>(dy), static_cast<const bf16*>(x), stats,
                 (const double*)ws, gamma, beta, eps, static_cast<bf16*>(dx), dbias_prev, lv, relu);
  return ut2_check_launch("groupnorm_bwd");
}

// stats: double[N*32*2] workspace owned by the caller (kept for backward).
extern "C" int ut2_groupnorm_relu_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y,
                                      double* stats, int N, int HW, int C, int G, int relu, void* stream) {
  return gn_fwd(x, gamma, beta, eps, y, stats, 1, &HW, N, C, G, relu, stream);
}

// ws: double[N*32*2] scratch; dgamma/dbeta are accumulated (+=) in fp32. dbias_prev (optional, float[256]) receives
// += the column sums of dx, i.e. the bias gradient of the convolution that feeds this GroupNorm.
extern "C" int ut2_groupnorm_relu_bwd(const void* dy, const void* x, const double* stats, const float* gamma,
                                      const float* beta, float eps, void* dx, float* dgamma, float* dbeta,
                                      float* dbias_prev, double* ws, int N, int HW, int C, int G, int relu,
                                      void* stream) {
  return gn_bwd(dy, x, stats, gamma, beta, eps, dx, dgamma, dbeta, dbias_prev, ws, 1, &HW, N, C, G, relu, stream);
}

// The same two operators over a level-major pyramid (hws: HOST array of H_l*W_l): GroupNorm statistics stay per
// (level, image, group) exactly as when fcos.py:338-376 applies the tower to each level separately; stats / ws are
// double[num_levels*N*32*2].
extern "C" int ut2_groupnorm_relu_levels_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y,
                                             double* stats, int num_levels, const int* hws, int N, int C, int G, int relu,
                                             void* stream) {
  return gn_fwd(x, gamma, beta, eps, y, stats, num_levels, hws, N, C, G, relu, stream);
}

extern "C" int ut2_groupnorm_relu_levels_bwd(const void* dy, const void* x, const double* stats, const float* gamma,
                                             const float* beta, float eps, void* dx, float* dgamma, float* dbeta,
                                             float* dbias_prev, double* ws, int num_levels, const int* hws, int N, int C,
                                             int G, int relu, void* stream) {
  return gn_bwd(dy, x, stats, gamma, beta, eps, dx, dgamma, dbeta, dbias_prev, ws, num_levels, hws, N, C, G, relu, stream);
}
