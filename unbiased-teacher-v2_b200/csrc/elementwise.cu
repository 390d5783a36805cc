// Bandwidth-bound NHWC bf16 helpers around the tensor-core convolutions: stem (7x7 s2 from uint8),
// max-pool, FPN top-down add, ReLU-backward masking, zero-stuffing for strided dgrad, bias-grad
// column sums and the bf16 weight packers. All are one coalesced pass over their operands.
//
// Reference call sites (all reached through Detectron2 / torch in the reference):
//   pixel normalisation + ImageList padding   ubteacher/modeling/one_stage_detector.py:88-90,165-167
//   BasicStem / max_pool / FPN top-down        ubteacher/modeling/backbone/fpn.py:59-78 -> [D2]
#include "ut2_internal.h"
#include <stdlib.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float bflo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bfhi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------------------ stem
// y[n, p, q, 0:64] = relu(scale * conv7x7_s2_p3(norm(x))[...] + shift), x uint8 CHW (BGR).
// One CTA: 8 x 32 output pixels, 256 threads, thread = one pixel x 64 channels (fp32 FMA).
constexpr int ST_TH = 8, ST_TW = 32;
constexpr int ST_IH = ST_TH * 2 + 5, ST_IW = ST_TW * 2 + 5;   // 21 x 69 input patch

__global__ void __launch_bounds__(256)
stem_conv_kernel(const uint8_t* __restrict__ img, int h, int w, const float* __restrict__ wgt /*[7][7][3][64]*/,
                 const float* __restrict__ scale, const float* __restrict__ shift, float m0, float m1,
                 float m2, float is0, float is1, float is2, bf16* __restrict__ out, int P, int Q) {
  extern __shared__ float sm[];
  float* sw = sm;                       // 147 * 64
  float* sx = sm + 147 * 64;            // 3 * ST_IH * ST_IW
  const int tid = threadIdx.x;
  for (int i = tid; i < 147 * 64; i += 256) sw[i] = wgt[i];
  const int p0 = blockIdx.y * ST_TH, q0 = blockIdx.x * ST_TW;
  const int ih0 = p0 * 2 - 3, iw0 = q0 * 2 - 3;
  const float mean[3] = {m0, m1, m2}, istd[3] = {is0, is1, is2};
  for (int i = tid; i < 3 * ST_IH * ST_IW; i += 256) {
    const int c = i / (ST_IH * ST_IW), r = (i / ST_IW) % ST_IH, col = i % ST_IW;
    const int ih = ih0 + r, iw = iw0 + col;
    float v = 0.f;   // conv padding and ImageList padding are both zeros *after* normalisation
    if (ih >= 0 && ih < h && iw >= 0 && iw < w)
      v = (static_cast<float>(img[(size_t)c * h * w + (size_t)ih * w + iw]) - mean[c]) * istd[c];
    sx[i] = v;
  }
  __syncthreads();
  const int tp = tid / ST_TW, tq = tid % ST_TW;
  const int p = p0 + tp, q = q0 + tq;
  float acc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 0.f;
  for (int r = 0; r < 7; ++r) {
    for (int s = 0; s < 7; ++s) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float xv = sx[c * ST_IH * ST_IW + (tp * 2 + r) * ST_IW + tq * 2 + s];
        const float4* wr = reinterpret_cast<const float4*>(sw + ((r * 7 + s) * 3 + c) * 64);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float4 w4 = wr[k];
          acc[4 * k + 0] = fmaf(xv, w4.x, acc[4 * k + 0]);
          acc[4 * k + 1] = fmaf(xv, w4.y, acc[4 * k + 1]);
          acc[4 * k + 2] = fmaf(xv, w4.z, acc[4 * k + 2]);
          acc[4 * k + 3] = fmaf(xv, w4.w, acc[4 * k + 3]);
        }
      }
    }
  }
  if (p < P && q < Q) {
    uint4* op = reinterpret_cast<uint4*>(out + ((size_t)p * Q + q) * 64);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        f[j] = fmaxf(fmaf(acc[8 * k + j], scale[8 * k + j], shift[8 * k + j]), 0.f);
      uint4 o;
      o.x = pack2(f[0], f[1]); o.y = pack2(f[2], f[3]); o.z = pack2(f[4], f[5]); o.w = pack2(f[6], f[7]);
      op[k] = o;
    }
  }
}

// ------------------------------------------------------------------------------ max-pool 3x3 s2 p1
__global__ void maxpool_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int N, int H, int W,
                               int C8, int P, int Q) {
  const size_t total = (size_t)N * P * Q * C8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = i % C8;
    size_t t = i / C8;
    const int q = t % Q; t /= Q;
    const int p = t % P;
    const int n = t / P;
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -3.0e38f;
    for (int dh = -1; dh <= 1; ++dh) {
      const int ih = p * 2 + dh;
      if (ih < 0 || ih >= H) continue;
      for (int dw = -1; dw <= 1; ++dw) {
        const int iw = q * 2 + dw;
        if (iw < 0 || iw >= W) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + (((size_t)n * H + ih) * W + iw) * C8 + c);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          m[2 * j] = fmaxf(m[2 * j], bflo(u[j]));
          m[2 * j + 1] = fmaxf(m[2 * j + 1], bfhi(u[j]));
        }
      }
    }
    uint4 o;
    o.x = pack2(m[0], m[1]); o.y = pack2(m[2], m[3]); o.z = pack2(m[4], m[5]); o.w = pack2(m[6], m[7]);
    reinterpret_cast<uint4*>(y)[i] = o;
  }
}

// ------------------------------------------------------------------------------ FPN top-down
// out[n,h,w,:] = lat[n,h,w,:] + top[n,h/2,w/2,:]   (nearest 2x upsample + sum)
__global__ void upsample_add_kernel(const bf16* __restrict__ lat, const bf16* __restrict__ top,
                                    bf16* __restrict__ out, int N, int H, int W, int C8) {
  const size_t total = (size_t)N * H * W * C8;
  const int Ht = H / 2, Wt = W / 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = i % C8;
    size_t t = i / C8;
    const int w = t % W; t /= W;
    const int h = t % H;
    const int n = t / H;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(lat) + i);
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(top) + (((size_t)n * Ht + h / 2) * Wt + w / 2) * C8 + c);
    const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = pack2(bflo(ua[j]) + bflo(ub[j]), bfhi(ua[j]) + bfhi(ub[j]));
    reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// backward of the nearest upsample: gtop[n,h,w,:] (+)= sum of the 2x2 children of g
__global__ void downsample_sum_kernel(const bf16* __restrict__ g, const bf16* __restrict__ addend,
                                      bf16* __restrict__ gtop, int N, int Ht, int Wt, int C8) {
  const size_t total = (size_t)N * Ht * Wt * C8;
  const int H = Ht * 2, W = Wt * 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = i % C8;
    size_t t = i / C8;
    const int w = t % Wt; t /= Wt;
    const int h = t % Ht;
    const int n = t / Ht;
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    if (addend) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(addend) + i);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[2 * j] = bflo(u[j]); s[2 * j + 1] = bfhi(u[j]); }
    }
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(g) + (((size_t)n * H + 2 * h + dh) * W + 2 * w + dw) * C8 + c);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { s[2 * j] += bflo(u[j]); s[2 * j + 1] += bfhi(u[j]); }
      }
    reinterpret_cast<uint4*>(gtop)[i] = make_uint4(pack2(s[0], s[1]), pack2(s[2], s[3]), pack2(s[4], s[5]), pack2(s[6], s[7]));
  }
}

// ------------------------------------------------------------------------------ relu backward (+ optional addend)
// g = (dy [+ dy2]) * (y > 0)
__global__ void relu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ dy2,
                                const bf16* __restrict__ y, bf16* __restrict__ g, size_t n8) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8;
       i += (size_t)gridDim.x * blockDim.x) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(dy) + i);
    const uint4 m = __ldg(reinterpret_cast<const uint4*>(y) + i);
    uint32_t ua[4] = {a.x, a.y, a.z, a.w};
    const uint32_t um[4] = {m.x, m.y, m.z, m.w};
    float f[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[2 * j] = bflo(ua[j]); f[2 * j + 1] = bfhi(ua[j]); }
    if (dy2) {
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(dy2) + i);
      const uint32_t ub[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { f[2 * j] += bflo(ub[j]); f[2 * j + 1] += bfhi(ub[j]); }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!(bflo(um[j]) > 0.f)) f[2 * j] = 0.f;
      if (!(bfhi(um[j]) > 0.f)) f[2 * j + 1] = 0.f;
    }
    reinterpret_cast<uint4*>(g)[i] = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
  }
}

// out = a + b (bf16, fp32 add) — gradient fan-in
__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ o, size_t n8) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8;
       i += (size_t)gridDim.x * blockDim.x) {
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(a) + i);
    const uint4 y = __ldg(reinterpret_cast<const uint4*>(b) + i);
    const uint32_t ux[4] = {x.x, x.y, x.z, x.w}, uy[4] = {y.x, y.y, y.z, y.w};
    uint32_t r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) r[j] = pack2(bflo(ux[j]) + bflo(uy[j]), bfhi(ux[j]) + bfhi(uy[j]));
    reinterpret_cast<uint4*>(o)[i] = make_uint4(r[0], r[1], r[2], r[3]);
  }
}

// ------------------------------------------------------------------------------ zero stuffing (stride-2 dgrad)
// out[n, 2p+oh, 2q+ow, :] = in[n, p, q, :], zeros elsewhere; out is [N, H, W, C].
__global__ void zero_stuff_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int N, int P, int Q,
                                  int H, int W, int C8, int oh, int ow) {
  const size_t total = (size_t)N * H * W * C8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = i % C8;
    size_t t = i / C8;
    const int w = t % W; t /= W;
    const int h = t % H;
    const int n = t / H;
    uint4 v = make_uint4(0, 0, 0, 0);
    const int hh = h - oh, ww = w - ow;
    if (hh >= 0 && ww >= 0 && !(hh & 1) && !(ww & 1) && (hh >> 1) < P && (ww >> 1) < Q)
      v = __ldg(reinterpret_cast<const uint4*>(in) + (((size_t)n * P + (hh >> 1)) * Q + (ww >> 1)) * C8 + c);
    reinterpret_cast<uint4*>(out)[i] = v;
  }
}

// ------------------------------------------------------------------------------ column sums (bias grad)
// db[c] += sum_m g[m, c]; g is [M, C] bf16, C % 8 == 0. One thread = one 16-byte vector (8 channels) of a row, so a
// warp reads whole 512-byte rows; per-thread fp32 partials over a strided set of rows, shared-memory reduce across
// the threads that own the same channel group, one atomicAdd per channel per block.
template <int UNR>
__global__ void __launch_bounds__(256)
colsum_kernel(const bf16* __restrict__ g, float* __restrict__ db, int M, int C8, int rows_per_block) {
  const int rpp = 256 / C8;                       // rows per pass
  const int cg = threadIdx.x % C8, ro = threadIdx.x / C8;
  const int m0 = blockIdx.x * rows_per_block;
  const int m1 = min(M, m0 + rows_per_block);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (ro < rpp) {
    for (int m = m0 + ro; m < m1; m += UNR * rpp) {
      uint4 v[UNR];
#pragma unroll
      for (int q = 0; q < UNR; ++q) {
        const int mm = m + q * rpp;
        v[q] = mm < m1 ? __ldg(reinterpret_cast<const uint4*>(g) + (size_t)mm * C8 + cg) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int q = 0; q < UNR; ++q) {
        const uint32_t u[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { s[2 * j] += bflo(u[j]); s[2 * j + 1] += bfhi(u[j]); }
      }
    }
  }
  __shared__ float red[256][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = s[j];
  __syncthreads();
  if (ro == 0) {
    for (int k = 1; k < rpp; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += red[k * C8 + cg][j];
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(db + cg * 8 + j, s[j]);
  }
}

// Batched column sums: up to 16 (gradient matrix, bias-gradient) pairs in ONE launch. The conv bias gradients of a backward
// pass (FPN laterals / outputs, P6 / P7, predictors, RPN head: a dozen matrices from 77 to 270 000 rows) cost more in launch
// tails than in bandwidth when they run one by one; here every block takes a row range of one matrix.
constexpr int CS_MAX = 16;
struct ColsumBatch {
  const bf16* g[CS_MAX];
  float* db[CS_MAX];
  int M[CS_MAX], C8[CS_MAX], rows[CS_MAX];
  int blk_off[CS_MAX + 1];
  int n;
};
__global__ void __launch_bounds__(256)
colsum_batched_kernel(const __grid_constant__ ColsumBatch b) {
  int e = 0;
#pragma unroll
  for (int i = 1; i < CS_MAX; ++i)
    if (i < b.n && (int)blockIdx.x >= b.blk_off[i]) e = i;
  const bf16* g = b.g[0];
  float* db = b.db[0];
  int M = b.M[0], C8 = b.C8[0], rows = b.rows[0], off = b.blk_off[0];
#pragma unroll
  for (int i = 1; i < CS_MAX; ++i)
    if (e == i) { g = b.g[i]; db = b.db[i]; M = b.M[i]; C8 = b.C8[i]; rows = b.rows[i]; off = b.blk_off[i]; }
  const int rpp = 256 / C8;
  const int cg = threadIdx.x % C8, ro = threadIdx.x / C8;
  const int m0 = ((int)blockIdx.x - off) * rows;
  const int m1 = min(M, m0 + rows);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (ro < rpp) {
    for (int m = m0 + ro; m < m1; m += 4 * rpp) {
      uint4 v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int mm = m + q * rpp;
        v[q] = mm < m1 ? __ldg(reinterpret_cast<const uint4*>(g) + (size_t)mm * C8 + cg) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t u[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { s[2 * j] += bflo(u[j]); s[2 * j + 1] += bfhi(u[j]); }
      }
    }
  }
  __shared__ float red[256][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = s[j];
  __syncthreads();
  if (ro == 0) {
    for (int k = 1; k < rpp; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += red[k * C8 + cg][j];
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(db + cg * 8 + j, s[j]);
  }
}

// ------------------------------------------------------------------------------ weight packing
// master fp32 weight, logical [Cout, Cin, R, S] stored channels-last (physical [Cout, R, S, Cin]) ->
//   wf  bf16 [Cout_pad, R, S, Cin]           forward B operand (rows >= Cout zeroed)
//   wt  bf16 [Cin_pad?, R, S, Cout_padT]     dgrad B operand: wt[c, r, s, n] = w[n, R-1-r, S-1-s, c]
__global__ void pack_weight_kernel(const float* __restrict__ w, bf16* __restrict__ wf, bf16* __restrict__ wt,
                                   int Cout, int Cin, int R, int S, int CoutT /*row length of wt*/) {
  const size_t total = (size_t)Cout * R * S * Cin;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = i % Cin;
    size_t t = i / Cin;
    const int s = t % S; t /= S;
    const int r = t % R;
    const int n = t / R;
    const bf16 v = __float2bfloat16_rn(w[i]);
    if (wf) wf[i] = v;
    if (wt) wt[(((size_t)c * R + (R - 1 - r)) * S + (S - 1 - s)) * CoutT + n] = v;
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn(x[i]);
}

// FrozenBN fold: scale = w * rsqrt(var + eps), shift = b - mean * scale   ([D2] FrozenBatchNorm2d)
__global__ void frozen_bn_fold_kernel(const float* w, const float* b, const float* mean, const float* var,
                                      float eps, float* scale, float* shift, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    const float sc = w[i] * rsqrtf(var[i] + eps);
    scale[i] = sc;
    shift[i] = b[i] - mean[i] * sc;
  }
}

inline int grid_for(size_t n, int block = 256) {
  size_t g = (n + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int ut2_stem_conv_u8(const void* img_chw, int h, int w, const float* wgt_rsck, const float* scale,
                                const float* shift, float m0, float m1, float m2, float s0, float s1,
                                float s2, void* out, int P, int Q, void* stream) {
  if (!img_chw || !wgt_rsck || !out) return ut2_fail(-1, "stem: null pointer");
  const int smem = (147 * 64 + 3 * ST_IH * ST_IW) * 4;
  static bool set = false;
  if (!set) {
    cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    set = true;
  }
  dim3 grid((Q + ST_TW - 1) / ST_TW, (P + ST_TH - 1) / ST_TH);
  stem_conv_kernel<<<grid, 256, smem, STREAM>>>(static_cast<const uint8_t*>(img_chw), h, w, wgt_rsck, scale, shift,
                                                m0, m1, m2, 1.f / s0, 1.f / s1, 1.f / s2,
                                                static_cast<bf16*>(out), P, Q);
  return ut2_check_launch("stem_conv");
}

extern "C" int ut2_maxpool3x3s2_nhwc(const void* x, void* y, int N, int H, int W, int C, void* stream) {
  if (C % 8) return ut2_fail(-2, "maxpool: C % 8 != 0");
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  const size_t total = (size_t)N * P * Q * (C / 8);
  maxpool_kernel<<<grid_for(total), 256, 0, STREAM>>>(static_cast<const bf16*>(x), static_cast<bf16*>(y), N, H, W, C / 8, P, Q);
  return ut2_check_launch("maxpool");
}

extern "C" int ut2_upsample2x_add_nhwc(const void* lat, const void* top, void* out, int N, int H, int W, int C,
                                       void* stream) {
  if (C % 8 || H % 2 || W % 2) return ut2_fail(-2, "upsample_add: need C%8==0 and even H, W");
  const size_t total = (size_t)N * H * W * (C / 8);
  upsample_add_kernel<<<grid_for(total), 256, 0, STREAM>>>(static_cast<const bf16*>(lat), static_cast<const bf16*>(top),
                                                           static_cast<bf16*>(out), N, H, W, C / 8);
  return ut2_check_launch("upsample_add");
}

extern "C" int ut2_downsample2x_sum_nhwc(const void* g, const void* addend, void* gtop, int N, int Ht, int Wt,
                                         int C, void* stream) {
  if (C % 8) return ut2_fail(-2, "downsample_sum: C % 8 != 0");
  const size_t total = (size_t)N * Ht * Wt * (C / 8);
  downsample_sum_kernel<<<grid_for(total), 256, 0, STREAM>>>(static_cast<const bf16*>(g), static_cast<const bf16*>(addend),
                                                             static_cast<bf16*>(gtop), N, Ht, Wt, C / 8);
  return ut2_check_launch("downsample_sum");
}

extern "C" int ut2_relu_bwd_bf16(const void* dy, const void* dy2, const void* y, void* g, long long n, void* stream) {
  if (n % 8) return ut2_fail(-2, "relu_bwd: n % 8 != 0");
  relu_bwd_kernel<<<grid_for(n / 8), 256, 0, STREAM>>>(static_cast<const bf16*>(dy), static_cast<const bf16*>(dy2),
                                                       static_cast<const bf16*>(y), static_cast<bf16*>(g), (size_t)n / 8);
  return ut2_check_launch("relu_bwd");
}

extern "C" int ut2_add_bf16(const void* a, const void* b, void* out, long long n, void* stream) {
  if (n % 8) return ut2_fail(-2, "add: n % 8 != 0");
  add_kernel<<<grid_for(n / 8), 256, 0, STREAM>>>(static_cast<const bf16*>(a), static_cast<const bf16*>(b),
                                                  static_cast<bf16*>(out), (size_t)n / 8);
  return ut2_check_launch("add");
}

extern "C" int ut2_zero_stuff_s2_nhwc(const void* in, void* out, int N, int P, int Q, int H, int W, int C, int oh,
                                      int ow, void* stream) {
  if (C % 8) return ut2_fail(-2, "zero_stuff: C % 8 != 0");
  const size_t total = (size_t)N * H * W * (C / 8);
  zero_stuff_kernel<<<grid_for(total), 256, 0, STREAM>>>(static_cast<const bf16*>(in), static_cast<bf16*>(out), N, P, Q, H,
                                                         W, C / 8, oh, ow);
  return ut2_check_launch("zero_stuff");
}

extern "C" int ut2_colsum_bf16(const void* g, float* db, int M, int C, void* stream) {
  if (C % 8 || C > 2048 || C <= 0) return ut2_fail(-2, "colsum: need C % 8 == 0 and C <= 2048");
  const int C8 = C / 8;
  // one block per SM, except for the >= 256 MB gradients of the R-CNN p2 level / level-major RPN conv, where more 16-byte
  // loads in flight pay for the extra per-block reductions and atomics (measured: 148 is better below, 4 x 148 above)
  static int mult = -1, unr = -1;
  if (mult < 0) { const char* e = getenv("UT2_COLSUM_MULT"); mult = e ? atoi(e) : 0; e = getenv("UT2_COLSUM_UNR"); unr = e ? atoi(e) : 4; }
  int blocks = mult > 0 ? 148 * mult : (M >= 500000 ? 148 * 4 : 148);
  int rows = (M + blocks - 1) / blocks;
  const int rpp = 256 / C8;
  if (rows < 8 * rpp) rows = 8 * rpp;
  blocks = (M + rows - 1) / rows;
  if (unr == 8) colsum_kernel<8><<<blocks, 256, 0, STREAM>>>(static_cast<const bf16*>(g), db, M, C8, rows);
  else colsum_kernel<4><<<blocks, 256, 0, STREAM>>>(static_cast<const bf16*>(g), db, M, C8, rows);
  return ut2_check_launch("colsum");
}

// gs / dbs / Ms / Cs: HOST arrays of n <= 16 device pointers / sizes (C % 8 == 0, C <= 256 so that a block covers whole rows).
extern "C" int ut2_colsum_bf16_batched(const void* const* gs, float* const* dbs, const int* Ms, const int* Cs, int n, void* stream) {
  if (n <= 0) return 0;
  if (n > CS_MAX || !gs || !dbs || !Ms || !Cs) return ut2_fail(-1, "colsum_batched: 1..16 entries");
  ColsumBatch b;
  long long total = 0;
  for (int i = 0; i < n; ++i) {
    if (Cs[i] % 8 || Cs[i] <= 0 || Cs[i] > 2048 || !gs[i] || !dbs[i]) return ut2_fail(-2, "colsum_batched: need C % 8 == 0, C <= 2048");
    total += (long long)Ms[i] * Cs[i];
  }
  // two blocks per SM in total for the FCOS-sized batches (a few hundred MB), up to 16 per SM for the R-CNN ones (the p2 level
  // and the level-major RPN gradient alone are 0.5 - 0.7 GB): about one block per MiB; at least 8 row passes per block
  long long nblk = total * 2 / (1 << 20);
  if (nblk < 148 * 2) nblk = 148 * 2;
  if (nblk > 148 * 16) nblk = 148 * 16;
  const long long per_block = (total + nblk - 1) / nblk;
  b.n = n;
  b.blk_off[0] = 0;
  for (int i = 0; i < CS_MAX; ++i) {
    const int j = i < n ? i : n - 1;
    b.g[i] = static_cast<const bf16*>(gs[j]); b.db[i] = dbs[j]; b.M[i] = Ms[j]; b.C8[i] = Cs[j] / 8;
    const int rpp = 256 / b.C8[i] > 0 ? 256 / b.C8[i] : 1;
    long long rows = per_block / Cs[j];
    if (rows < 8 * rpp) rows = 8 * rpp;
    b.rows[i] = (int)rows;
    b.blk_off[i + 1] = b.blk_off[i] + (i < n ? (int)((Ms[j] + rows - 1) / rows) : 0);
  }
  colsum_batched_kernel<<<b.blk_off[n], 256, 0, STREAM>>>(b);
  return ut2_check_launch("colsum_batched");
}

extern "C" int ut2_pack_conv_weight(const float* w, void* wf, void* wt, int Cout, int Cin, int R, int S, int CoutT,
                                    void* stream) {
  const size_t total = (size_t)Cout * R * S * Cin;
  pack_weight_kernel<<<grid_for(total), 256, 0, STREAM>>>(w, static_cast<bf16*>(wf), static_cast<bf16*>(wt), Cout, Cin,
                                                          R, S, CoutT);
  return ut2_check_launch("pack_weight");
}

extern "C" int ut2_cast_f32_bf16(const float* x, void* y, long long n, void* stream) {
  cast_f32_bf16_kernel<<<grid_for(n), 256, 0, STREAM>>>(x, static_cast<bf16*>(y), (size_t)n);
  return ut2_check_launch("cast");
}

extern "C" int ut2_frozen_bn_fold(const float* w, const float* b, const float* mean, const float* var, float eps,
                                  float* scale, float* shift, int C, void* stream) {
  frozen_bn_fold_kernel<<<(C + 127) / 128, 128, 0, STREAM>>>(w, b, mean, var, eps, scale, shift, C);
  return ut2_check_launch("frozen_bn_fold");
}
