// Stand-alone loss operators with the reference's module interface (ubteacher/layers/iou_loss.py:23-76, kl_loss.py:17-105):
// IOULoss (iou / linear_iou / giou), NLLoss, KLLoss on [P, 4] rows. The training step computes the same terms inside the fused
// FCOS loss kernels (csrc/fcos_loss.cu); these are the operators behind ubteacher.layers.{IOULoss, NLLoss, KLLoss}. One thread
// per row, forward value and input gradients in the same pass (the autograd wrapper scales the stored gradients).
#include "ut2_internal.h"
#include <math.h>

namespace {

__device__ __forceinline__ void warp_add_double(double v, double* dst) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(dst, v);
}
// d min(a, b) / da and d max(a, b) / da with torch's even split on ties
__device__ __forceinline__ float dmin_a(float a, float b) { return a < b ? 1.f : (a == b ? 0.5f : 0.f); }
__device__ __forceinline__ float dmax_a(float a, float b) { return a > b ? 1.f : (a == b ? 0.5f : 0.f); }

// type: 0 iou (-log), 1 linear_iou, 2 giou. pred / target: (l, t, r, b) distances.
__global__ void iou_loss_kernel(const float4* __restrict__ pred, const float4* __restrict__ tgt, const float* __restrict__ weight, int P,
                                int type, double* __restrict__ acc, float4* __restrict__ dpred) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double loss = 0.0;
  if (i < P) {
    const float4 p = pred[i], t = tgt[i];
    const float w = weight ? weight[i] : 1.f;
    const float At = (t.x + t.z) * (t.y + t.w), Ap = (p.x + p.z) * (p.y + p.w);
    const float wi = fminf(p.x, t.x) + fminf(p.z, t.z), hi = fminf(p.w, t.w) + fminf(p.y, t.y);
    const float gw = fmaxf(p.x, t.x) + fmaxf(p.z, t.z), gh = fmaxf(p.w, t.w) + fmaxf(p.y, t.y);
    const float C = gw * gh, I = wi * hi, U = At + Ap - I;
    const float iou = (I + 1.f) / (U + 1.f);
    const float giou = iou - (C - U) / C;
    const float l = type == 0 ? -logf(iou) : (type == 1 ? 1.f - iou : 1.f - giou);
    loss = (double)(l * w);
    if (dpred) {
      // order (l, t, r, b) = (x, y, z, w)
      const float dAp[4] = {p.y + p.w, p.x + p.z, p.y + p.w, p.x + p.z};
      const float dI[4] = {hi * dmin_a(p.x, t.x), wi * dmin_a(p.y, t.y), hi * dmin_a(p.z, t.z), wi * dmin_a(p.w, t.w)};
      const float dC[4] = {gh * dmax_a(p.x, t.x), gw * dmax_a(p.y, t.y), gh * dmax_a(p.z, t.z), gw * dmax_a(p.w, t.w)};
      float g[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float dU = dAp[k] - dI[k];
        const float diou = (dI[k] * (U + 1.f) - (I + 1.f) * dU) / ((U + 1.f) * (U + 1.f));
        const float dgiou = diou + (dU * C - U * dC[k]) / (C * C);
        g[k] = w * (type == 0 ? -diou / iou : (type == 1 ? -diou : -dgiou));
      }
      dpred[i] = make_float4(g[0], g[1], g[2], g[3]);
    }
  }
  warp_add_double(loss, acc);
}

// NLLoss (kl_loss.py:75-105): mean_i [ sum_4 ((t - mu)^2 / (2 sigma^2) + 0.5 log sigma^2) + 2 log(2 pi) ] * iou_weight_i, sigma = sigmoid(std)
__global__ void nl_loss_kernel(const float4* __restrict__ mu, const float4* __restrict__ sd, const float4* __restrict__ tgt,
                               const float* __restrict__ iouw, int P, double* __restrict__ acc, float4* __restrict__ dmu,
                               float4* __restrict__ dsd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double loss = 0.0;
  if (i < P) {
    const float4 m4 = mu[i], s4 = sd[i], t4 = tgt[i];
    const float m[4] = {m4.x, m4.y, m4.z, m4.w}, s[4] = {s4.x, s4.y, s4.z, s4.w}, t[4] = {t4.x, t4.y, t4.z, t4.w};
    const float w = iouw[i], invP = 1.f / (float)P;
    float sum = 2.f * logf(2.f * 3.14159265358979323846f), gm[4], gs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float sg = 1.f / (1.f + expf(-s[k]));
      const float sq = sg * sg, d = t[k] - m[k];
      sum += d * d / (2.f * sq) + 0.5f * logf(sq);
      gm[k] = -d / sq * w * invP;
      gs[k] = (-d * d / (sq * sg) + 1.f / sg) * sg * (1.f - sg) * w * invP;
    }
    loss = (double)(sum * w * invP);
    if (dmu) dmu[i] = make_float4(gm[0], gm[1], gm[2], gm[3]);
    if (dsd) dsd[i] = make_float4(gs[0], gs[1], gs[2], gs[3]);
  }
  warp_add_double(loss, acc);
}

// KLLoss (kl_loss.py:17-66), beta >= 1e-5: loss = exp(-std) * smooth_l1(input - target; beta) + 0.5 * std, then
// method 0 weight_ctr_sum: sum_i w_i sum_4; 1 weight_ctr_mean: the same / loss_denorm; 2 sum; 3 mean (over 4P elements)
__global__ void kl_loss_kernel(const float4* __restrict__ x, const float4* __restrict__ sd, const float4* __restrict__ tgt,
                               const float* __restrict__ weight, int P, float beta, float scale, double* __restrict__ acc,
                               float4* __restrict__ dx, float4* __restrict__ dsd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double loss = 0.0;
  if (i < P) {
    const float4 x4 = x[i], s4 = sd[i], t4 = tgt[i];
    const float xv[4] = {x4.x, x4.y, x4.z, x4.w}, s[4] = {s4.x, s4.y, s4.z, s4.w}, t[4] = {t4.x, t4.y, t4.z, t4.w};
    const float w = (weight ? weight[i] : 1.f) * scale;
    float sum = 0.f, gx[4], gs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = xv[k] - t[k], n = fabsf(d), e = expf(-s[k]);
      const bool quad = n < beta;
      const float l1 = quad ? 0.5f * n * n / beta : n - 0.5f * beta;
      const float dl = quad ? d / beta : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
      sum += e * l1 + 0.5f * s[k];
      gx[k] = e * dl * w;
      gs[k] = (-e * l1 + 0.5f) * w;
    }
    loss = (double)(sum * w);
    if (dx) dx[i] = make_float4(gx[0], gx[1], gx[2], gx[3]);
    if (dsd) dsd[i] = make_float4(gs[0], gs[1], gs[2], gs[3]);
  }
  warp_add_double(loss, acc);
}

__global__ void finish_scalar_kernel(const double* __restrict__ acc, float* __restrict__ out) { out[0] = (float)acc[0]; }
__global__ void scale_f32_kernel(const float* __restrict__ x, const float* __restrict__ s, float* __restrict__ y, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) y[i] = x[i] * s[0];
}
}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

// pred / target [P, 4] f32 (l, t, r, b), weight [P] or NULL; loss: float[1]; dpred [P, 4] or NULL (d loss / d pred); acc: double[1] scratch
extern "C" int ut2_iou_loss(const float* pred, const float* target, const float* weight, int P, int type, double* acc, float* loss,
                            float* dpred, void* stream) {
  if (!acc || !loss || (P > 0 && (!pred || !target))) return ut2_fail(-1, "iou_loss: null pointer");
  if (type < 0 || type > 2) return ut2_fail(-2, "iou_loss: type must be 0 (iou), 1 (linear_iou) or 2 (giou)");
  cudaMemsetAsync(acc, 0, sizeof(double), STREAM);
  if (P > 0)
    iou_loss_kernel<<<(P + 255) / 256, 256, 0, STREAM>>>(reinterpret_cast<const float4*>(pred), reinterpret_cast<const float4*>(target), weight,
                                                         P, type, acc, reinterpret_cast<float4*>(dpred));
  finish_scalar_kernel<<<1, 1, 0, STREAM>>>(acc, loss);
  return ut2_check_launch("iou_loss");
}

extern "C" int ut2_nl_loss(const float* mean, const float* std, const float* target, const float* iou_weight, int P, double* acc,
                           float* loss, float* dmean, float* dstd, void* stream) {
  if (!acc || !loss || P <= 0 || !mean || !std || !target || !iou_weight) return ut2_fail(-1, "nl_loss: null pointer / empty input");
  cudaMemsetAsync(acc, 0, sizeof(double), STREAM);
  nl_loss_kernel<<<(P + 255) / 256, 256, 0, STREAM>>>(reinterpret_cast<const float4*>(mean), reinterpret_cast<const float4*>(std),
                                                      reinterpret_cast<const float4*>(target), iou_weight, P, acc,
                                                      reinterpret_cast<float4*>(dmean), reinterpret_cast<float4*>(dstd));
  finish_scalar_kernel<<<1, 1, 0, STREAM>>>(acc, loss);
  return ut2_check_launch("nl_loss");
}

// method: 0 weight_ctr_sum, 1 weight_ctr_mean (divides by loss_denorm), 2 sum, 3 mean
extern "C" int ut2_kl_loss(const float* input, const float* std, const float* target, const float* weight, int P, float beta, int method,
                           float loss_denorm, double* acc, float* loss, float* dinput, float* dstd, void* stream) {
  if (!acc || !loss || P <= 0 || !input || !std || !target) return ut2_fail(-1, "kl_loss: null pointer / empty input");
  if (beta < 1e-5f) return ut2_fail(-2, "kl_loss: beta < 1e-5 (the reference returns None on that branch)");
  if (method < 0 || method > 3 || (method < 2 && !weight)) return ut2_fail(-2, "kl_loss: bad method / missing weight");
  const float scale = method == 1 ? 1.f / loss_denorm : (method == 3 ? 1.f / (4.f * (float)P) : 1.f);
  cudaMemsetAsync(acc, 0, sizeof(double), STREAM);
  kl_loss_kernel<<<(P + 255) / 256, 256, 0, STREAM>>>(reinterpret_cast<const float4*>(input), reinterpret_cast<const float4*>(std),
                                                      reinterpret_cast<const float4*>(target), method < 2 ? weight : nullptr, P, beta, scale,
                                                      acc, reinterpret_cast<float4*>(dinput), reinterpret_cast<float4*>(dstd));
  finish_scalar_kernel<<<1, 1, 0, STREAM>>>(acc, loss);
  return ut2_check_launch("kl_loss");
}

// y = x * s[0] (s on the device): scales stored gradients by the incoming scalar gradient without a host round trip
extern "C" int ut2_scale_f32(const float* x, const float* s, float* y, long long n, void* stream) {
  if (n <= 0) return 0;
  scale_f32_kernel<<<(int)((n + 255) / 256), 256, 0, STREAM>>>(x, s, y, n);
  return ut2_check_launch("scale_f32");
}
