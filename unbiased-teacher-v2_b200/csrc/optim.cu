// Flat-arena optimiser kernels: EMA teacher update, SGD-momentum step (+ grad zeroing) and the
// table-driven bf16 weight packer that refreshes every conv's tensor-core operands in one launch.
//
// Reference: ubteacher/engine/trainer.py:468-486 (_update_teacher_model: new = s*(1-k) + t*k over the whole
// state_dict), :422-429 (zero_grad / backward / optimizer.step with [D2]-built torch.optim.SGD).
// Algorithmic bytes: EMA 12 B/element (read s, read t, write t); SGD 20 B/element (p, g, buf read; p, buf
// write) + 4 B when the gradient is zeroed in the same pass.
#include "ut2_internal.h"
#include <cuda_bf16.h>
#include <stdint.h>

namespace {
typedef __nv_bfloat16 bf16;

// bit-exact with torch: both products rounded to fp32, then one rounded add (no FMA contraction)
__global__ void __launch_bounds__(256)
ema_kernel(const float4* __restrict__ s, float4* __restrict__ t, size_t n4, const float* __restrict__ s_tail,
           float* __restrict__ t_tail, int tail, float a, float b) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 sv = __ldg(s + i);
    float4 tv = t[i];
    tv.x = __fadd_rn(__fmul_rn(sv.x, a), __fmul_rn(tv.x, b));
    tv.y = __fadd_rn(__fmul_rn(sv.y, a), __fmul_rn(tv.y, b));
    tv.z = __fadd_rn(__fmul_rn(sv.z, a), __fmul_rn(tv.z, b));
    tv.w = __fadd_rn(__fmul_rn(sv.w, a), __fmul_rn(tv.w, b));
    t[i] = tv;
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail)
    t_tail[threadIdx.x] = __fadd_rn(__fmul_rn(s_tail[threadIdx.x], a), __fmul_rn(t_tail[threadIdx.x], b));
}

// torch.optim.SGD(momentum, weight_decay, dampening=0, nesterov=False):
//   g' = g + wd*p ; buf = first ? g' : mom*buf + g' ; p -= lr*buf ; (optionally g = 0)
__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ buf, size_t n, float lr,
           const float* __restrict__ lr_dev, float mom, float wd, int first, int zero_grad, float gscale) {
  if (lr_dev) lr = *lr_dev;      // device-resident learning rate: the step can live in a replayed CUDA graph
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float pv = p[i];
    const float gv = g[i] * gscale + wd * pv;
    const float bv = first ? gv : mom * buf[i] + gv;
    buf[i] = bv;
    p[i] = pv - lr * bv;
    if (zero_grad) g[i] = 0.f;
  }
}

struct PackDesc {          // one conv weight: master fp32 [Cout, R, S, Cin] (channels-last physical layout)
  long long src;           // float offset in the parameter arena
  long long wf;            // bf16 offset of the forward operand [rows, R, S, Cin] in the pack arena, or -1
  long long wt;            // bf16 offset of the dgrad operand [Cin, R, S, CoutT] (flipped taps), or -1
  long long begin;         // prefix sum of TILE counts: R * S * ceil(Cout / 64) * ceil(Cin / 32) tiles per descriptor
  long long scale;         // float offset of a per-Cout scale folded into the packed weights (FrozenBN), or -1
  int Cout, Cin, R, S, CoutT, n_off;   // n_off: column offset inside wt rows (fused predictors)
};

// One CTA per (descriptor, tap, 64 output channels x 32 input channels) tile: the fp32 master tile is read along Cin
// (128-byte rows), scaled, rounded to bf16 and written straight to the forward operand (same order); the dgrad operand
// is its transpose with flipped taps, written along Cout from a padded shared-memory tile so both sides coalesce.
constexpr int PACK_TN = 64, PACK_TC = 32;

__global__ void __launch_bounds__(256)
pack_tiles_kernel(const PackDesc* __restrict__ descs, int num, long long total_tiles, const float* __restrict__ arena,
                  const float* __restrict__ scales, bf16* __restrict__ packed, int write_dgrad) {
  __shared__ float tile[PACK_TN][PACK_TC + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int nl = threadIdx.x & 63, cb = threadIdx.x >> 6;
  for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    int lo = 0, hi = num - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (descs[mid].begin <= t) lo = mid; else hi = mid - 1;
    }
    const PackDesc d = descs[lo];
    int e = (int)(t - d.begin);
    const int ct = (d.Cin + PACK_TC - 1) / PACK_TC, nt = (d.Cout + PACK_TN - 1) / PACK_TN;
    const int ci = e % ct; e /= ct;
    const int ni = e % nt; e /= nt;
    const int s = e % d.S, r = e / d.S;
    const bool dg = write_dgrad && d.wt >= 0;
    if ((d.Cin & 1) == 0 && (d.src & 1) == 0 && (d.wf & 1) == 0) {
      // even Cin (every convolution but ragged test shapes): two channels per thread — 8-byte loads, bf16x2 stores
      const int tx2 = threadIdx.x & 15, ty2 = threadIdx.x >> 4;
      const int c = ci * PACK_TC + 2 * tx2;
#pragma unroll
      for (int k = 0; k < PACK_TN / 16; ++k) {
        const int n = ni * PACK_TN + ty2 + 16 * k;
        if (n < d.Cout && c < d.Cin) {
          const long long idx = (((long long)n * d.R + r) * d.S + s) * d.Cin + c;
          float2 wv = __ldg(reinterpret_cast<const float2*>(arena + d.src + idx));
          if (d.scale >= 0) { const float sc = __ldg(scales + d.scale + n); wv.x *= sc; wv.y *= sc; }
          const __nv_bfloat162 v = __floats2bfloat162_rn(wv.x, wv.y);
          if (d.wf >= 0) *reinterpret_cast<__nv_bfloat162*>(packed + d.wf + idx) = v;
          tile[ty2 + 16 * k][2 * tx2] = __low2float(v);
          tile[ty2 + 16 * k][2 * tx2 + 1] = __high2float(v);
        }
      }
    } else {
      const int c = ci * PACK_TC + tx;
#pragma unroll
      for (int k = 0; k < PACK_TN / 8; ++k) {
        const int n = ni * PACK_TN + ty + 8 * k;
        if (n < d.Cout && c < d.Cin) {
          const long long idx = (((long long)n * d.R + r) * d.S + s) * d.Cin + c;
          float wv = __ldg(arena + d.src + idx);
          if (d.scale >= 0) wv *= __ldg(scales + d.scale + n);
          const bf16 v = __float2bfloat16_rn(wv);
          if (d.wf >= 0) packed[d.wf + idx] = v;
          tile[ty + 8 * k][tx] = __bfloat162float(v);
        }
      }
    }
    if (dg) {
      __syncthreads();
      if (((d.CoutT | d.n_off | d.wt) & 1) == 0) {
        // two output channels per thread along the transposed operand's rows: bf16x2 stores
        const int np = threadIdx.x & 31, cb8 = threadIdx.x >> 5;          // 32 channel pairs x 8 input channels per pass
        const int n = ni * PACK_TN + 2 * np;
#pragma unroll
        for (int k = 0; k < PACK_TC / 8; ++k) {
          const int cl = cb8 + 8 * k, cc = ci * PACK_TC + cl;
          if (n < d.Cout && cc < d.Cin) {
            bf16* dst = packed + d.wt + (((long long)cc * d.R + (d.R - 1 - r)) * d.S + (d.S - 1 - s)) * d.CoutT + n + d.n_off;
            if (n + 1 < d.Cout) *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(tile[2 * np][cl], tile[2 * np + 1][cl]);
            else *dst = __float2bfloat16_rn(tile[2 * np][cl]);
          }
        }
      } else {
        const int n = ni * PACK_TN + nl;
#pragma unroll
        for (int k = 0; k < PACK_TC / 4; ++k) {
          const int cl = cb + 4 * k, cc = ci * PACK_TC + cl;
          if (n < d.Cout && cc < d.Cin)
            packed[d.wt + (((long long)cc * d.R + (d.R - 1 - r)) * d.S + (d.S - 1 - s)) * d.CoutT + n + d.n_off] =
                __float2bfloat16_rn(tile[nl][cl]);
        }
      }
      __syncthreads();
    }
  }
}

inline int grid_for(size_t n) {
  size_t g = (n + 255) / 256;
  const size_t cap = 148 * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}
}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int ut2_ema_update(const float* student, float* teacher, long long n, double keep_rate, void* stream) {
  if (n <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(student) | reinterpret_cast<uintptr_t>(teacher)) & 15)
    return ut2_fail(-2, "ema: arenas must be 16-byte aligned");
  const float a = (float)(1.0 - keep_rate), b = (float)keep_rate;
  const size_t n4 = (size_t)n / 4;
  const int tail = (int)(n - (long long)n4 * 4);
  ema_kernel<<<grid_for(n4), 256, 0, STREAM>>>(reinterpret_cast<const float4*>(student), reinterpret_cast<float4*>(teacher),
                                               n4, student + n4 * 4, teacher + n4 * 4, tail, a, b);
  return ut2_check_launch("ema_update");
}

extern "C" int ut2_sgd_step(float* p, float* g, float* buf, long long n, float lr, const float* lr_dev, float momentum,
                            float weight_decay, int first_step, int zero_grad, float grad_scale, void* stream) {
  if (n <= 0) return 0;
  sgd_kernel<<<grid_for(n), 256, 0, STREAM>>>(p, g, buf, (size_t)n, lr, lr_dev, momentum, weight_decay, first_step,
                                              zero_grad, grad_scale);
  return ut2_check_launch("sgd_step");
}

// descs: device array of `num` 64-byte records {int64 src, wf, wt, begin, scale; int32 Cout, Cin, R, S, CoutT, n_off};
// begin = prefix sum of R * S * ceil(Cout / 64) * ceil(Cin / 32) (tiles), total_tiles = its end. write_dgrad = 0 skips the
// transposed operands (an inference-only replica: the EMA teacher).
extern "C" int ut2_pack_conv_weights_batched(const void* descs, int num, long long total_tiles, const float* arena,
                                             const float* scales, void* packed, int write_dgrad, void* stream) {
  if (num <= 0 || total_tiles <= 0) return 0;
  const long long cap = 148 * 8;
  pack_tiles_kernel<<<(int)(total_tiles < cap ? total_tiles : cap), 256, 0, STREAM>>>(
      static_cast<const PackDesc*>(descs), num, total_tiles, arena, scales, static_cast<bf16*>(packed), write_dgrad);
  return ut2_check_launch("pack_conv_weights_batched");
}
