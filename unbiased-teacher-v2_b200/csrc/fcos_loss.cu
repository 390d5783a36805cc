// FCOS target assignment and the supervised / pseudo-label losses of Unbiased Teacher v2, forward and
// backward, as coalesced single-pass kernels with device-side normalisers (no host synchronisation).
//
// Reference (paths under /root/reference/ubteacher):
//   targets      modeling/fcos/fcos_outputs.py:649-698 (_get_ground_truth), :772-906
//   sup. losses  modeling/fcos/fcos_outputs.py:307-444 (fcos_losses)
//   pseudo       modeling/fcos/fcos_outputs.py:492-631 (fcos_pseudo_losses, class_loss)
//   pieces       fcos_outputs.py:44-129 (Integral, ctrness/iou targets), layers/iou_loss.py:23-76,
//                layers/kl_loss.py:75-105 (NLLoss), [fvcore] sigmoid_focal_loss_jit
//
// Data layout: every per-location tensor is "level-major": position p = level_off[l]*N + img*HW_l + hw,
// which is exactly the memory of the per-level NHWC head outputs laid back to back
// (cls_out [P, 80] bf16; box_out [P, 80] bf16 = 68 distribution logits | 4 std | 1 centerness | 7 pad).
// The learnable per-level Scale (fcos.py:22-29,367) is applied here, not in the conv epilogue.
#include "ut2_internal.h"
#include <cuda_bf16.h>
#include <stdint.h>

namespace {
typedef __nv_bfloat16 bf16;
constexpr int MAXL = 8;
constexpr float BG_AREA = 100000000.0f;   // INF in the reference

struct Levels {
  int num;
  int H[MAXL], W[MAXL], stride[MAXL];
  int off[MAXL + 1];       // prefix of H*W
  float lo[MAXL], hi[MAXL];
};

__device__ __forceinline__ void locate(const Levels& lv, int N, long long p, int& l, int& img, int& hw) {
  l = 0;
#pragma unroll
  for (int i = 1; i < MAXL; ++i)
    if (i < lv.num && p >= (long long)lv.off[i] * N) l = i;
  const long long r = p - (long long)lv.off[l] * N;
  const int HW = lv.H[l] * lv.W[l];
  img = (int)(r / HW);
  hw = (int)(r - (long long)img * HW);
}

__device__ __forceinline__ float ctr_target(const float (&t)[4]) {
  const float lr = __fdiv_rn(fminf(t[0], t[2]), fmaxf(t[0], t[2]));
  const float tb = __fdiv_rn(fminf(t[1], t[3]), fmaxf(t[1], t[3]));
  return __fsqrt_rn(__fmul_rn(lr, tb));
}

// ------------------------------------------------------------------------------------ targets
__global__ void __launch_bounds__(256)
assign_targets_kernel(Levels lv, int N, int G, const float* __restrict__ boxes, const long long* __restrict__ classes,
                      const int* __restrict__ counts, const float* __restrict__ bvar, int num_classes,
                      float center_radius, int ignore_near, long long* __restrict__ labels, long long* __restrict__ tinds, float* __restrict__ reg_t,
                      float* __restrict__ bv_out, uint8_t* __restrict__ keep, float* __restrict__ norm) {
  const long long P = (long long)lv.off[lv.num] * N;
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float my_pos = 0.f, my_ctr = 0.f;
  if (p < P) {
    int l, img, hw;
    locate(lv, N, p, l, img, hw);
    const int h = hw / lv.W[l], w = hw - h * lv.W[l];
    const float x = (float)(w * lv.stride[l]) + (float)(lv.stride[l] / 2);
    const float y = (float)(h * lv.stride[l]) + (float)(lv.stride[l] / 2);
    const int n = counts[img];
    long long lab = num_classes, ti = -1;
    float t[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    uint8_t kp = 0;
    if (n > 0) {
      kp = 1;
      int best = 0;
      float best_area = BG_AREA;
      const float* b = boxes + (size_t)img * G * 4;
      // CENTER_SAMPLE (get_sample_region, fcos_outputs.py:700-770): a location is positive for a box only inside the box's
      // centre region, [centre -/+ stride * radius] clipped to the box. The reference returns an all-false mask when the
      // FIRST box of the image has centre x == 0 (its "no gt" test, :736).
      const bool center = center_radius > 0.f;
      const float sr = __fmul_rn((float)lv.stride[l], center_radius);
      const bool center_none = center && __fmul_rn(b[0] + b[2], 0.5f) == 0.f;
      bool any_inside = false, any_region = false;
      for (int j = 0; j < n; ++j) {
        const float x1 = b[4 * j], y1 = b[4 * j + 1], x2 = b[4 * j + 2], y2 = b[4 * j + 3];
        const float dl = x - x1, dt = y - y1, dr = x2 - x, db = y2 - y;
        const float mn = fminf(fminf(dl, dt), fminf(dr, db));
        const float mx = fmaxf(fmaxf(dl, dt), fmaxf(dr, db));
        float a = __fmul_rn(x2 - x1, y2 - y1);
        bool in = mn > 0.f;
        any_inside |= in;
        if (center) {
          const float cx = __fmul_rn(x1 + x2, 0.5f), cy = __fmul_rn(y1 + y2, 0.5f);
          const float xmin = cx - sr, ymin = cy - sr, xmax = cx + sr, ymax = cy + sr;
          const float rx1 = xmin > x1 ? xmin : x1, ry1 = ymin > y1 ? ymin : y1;
          const float rx2 = xmax > x2 ? x2 : xmax, ry2 = ymax > y2 ? y2 : ymax;
          in = !center_none && fminf(fminf(x - rx1, y - ry1), fminf(rx2 - x, ry2 - y)) > 0.f;
        }
        any_region |= in;
        if (!in) a = BG_AREA;
        if (!(mx >= lv.lo[l] && mx <= lv.hi[l])) a = BG_AREA;
        if (a < best_area) { best_area = a; best = j; }     // first minimum wins
      }
      if (ignore_near) kp = (!any_inside || any_region) ? 1 : 0;      // :841-848: drop locations inside a box but off every centre region
      int prefix = 0;
      for (int i = 0; i < img; ++i) prefix += counts[i];
      ti = (long long)best + prefix;
      const float x1 = b[4 * best], y1 = b[4 * best + 1], x2 = b[4 * best + 2], y2 = b[4 * best + 3];
      const float s = (float)lv.stride[l];
      t[0] = __fdiv_rn(x - x1, s); t[1] = __fdiv_rn(y - y1, s);
      t[2] = __fdiv_rn(x2 - x, s); t[3] = __fdiv_rn(y2 - y, s);
      if (best_area == BG_AREA) {
        bv[0] = bv[1] = bv[2] = bv[3] = 99999.0f;
      } else {
        lab = classes[(size_t)img * G + best];
        if (bvar) {
#pragma unroll
          for (int k = 0; k < 4; ++k) bv[k] = bvar[((size_t)img * G + best) * 4 + k];
        }
        my_pos = 1.f;
        my_ctr = ctr_target(t);
      }
    }
    labels[p] = lab;
    tinds[p] = ti;
    keep[p] = kp;
    reinterpret_cast<float4*>(reg_t)[p] = make_float4(t[0], t[1], t[2], t[3]);
    reinterpret_cast<float4*>(bv_out)[p] = make_float4(bv[0], bv[1], bv[2], bv[3]);
  }
  // block reduction of (num_pos, sum ctrness)
  __shared__ float r0[8], r1[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    my_pos += __shfl_xor_sync(0xffffffffu, my_pos, o);
    my_ctr += __shfl_xor_sync(0xffffffffu, my_ctr, o);
  }
  if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = my_pos; r1[threadIdx.x >> 5] = my_ctr; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int i = 0; i < 8; ++i) { a += r0[i]; c += r1[i]; }
    if (a != 0.f) { atomicAdd(norm, a); atomicAdd(norm + 1, c); }
  }
}

// ------------------------------------------------------------------------------------ focal
__device__ __forceinline__ float focal_term(float x, bool t, float alpha, float gamma, float* dldx) {
  // z = x for the target class, -x otherwise; p_t = sigmoid(z); loss = a_t (1-p_t)^gamma (-log p_t)
  const float z = t ? x : -x;
  const float e = expf(-fabsf(z));
  const float logpt = fminf(z, 0.f) - log1pf(e);            // log sigmoid(z)
  const float pt = (z >= 0.f) ? 1.f / (1.f + e) : e / (1.f + e);
  const float om = 1.f - pt;
  const float a_t = alpha >= 0.f ? (t ? alpha : 1.f - alpha) : 1.f;
  const float mod = (gamma == 2.f) ? om * om : powf(om, gamma);
  if (dldx) {
    const float dz = a_t * mod * (gamma * pt * logpt - om);
    *dldx = t ? dz : -dz;
  }
  return -a_t * mod * logpt;
}

// sum over all P x C logits; acc[0] += sum
__global__ void __launch_bounds__(256)
focal_fwd_kernel(const bf16* __restrict__ logits, int ld, int C, const long long* __restrict__ labels,
                 const uint8_t* __restrict__ keep, long long P, float alpha, float gamma, double* __restrict__ acc) {
  const long long total = P * (C / 2);
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / (C / 2);
    const int c = (int)(i - p * (C / 2)) * 2;
    if (keep && !keep[p]) continue;   // fcos_outputs.py:310 — images without GT are dropped from the labeled loss
    const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(logits + p * ld + c));
    const int lab = (int)labels[p];
    s += focal_term(__uint_as_float(u << 16), lab == c, alpha, gamma, nullptr);
    s += focal_term(__uint_as_float(u & 0xFFFF0000u), lab == c + 1, alpha, gamma, nullptr);
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = 0.0;
    for (int i = 0; i < 8; ++i) d += red[i];
    atomicAdd(acc, d);
  }
}

// dlogits = coef * dfocal/dx, coef = gout / max(norm[0] / world, 1)  (0 when zero_if_nopos and no positives)
__global__ void __launch_bounds__(256)
focal_bwd_kernel(const bf16* __restrict__ logits, int ld, int C, const long long* __restrict__ labels,
                 const uint8_t* __restrict__ keep, long long P, float alpha, float gamma,
                 const float* __restrict__ norm, float world, const float* __restrict__ gout,
                 const double* __restrict__ acc, int zero_if_nopos, bf16* __restrict__ dlogits) {
  float coef = gout[0] / fmaxf(norm[0] / world, 1.0f);
  if (zero_if_nopos && acc[6] == 0.0) coef = 0.f;
  const long long total = P * (C / 2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / (C / 2);
    const int c = (int)(i - p * (C / 2)) * 2;
    const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(logits + p * ld + c));
    const int lab = (int)labels[p];
    float d0, d1;
    focal_term(__uint_as_float(u << 16), lab == c, alpha, gamma, &d0);
    focal_term(__uint_as_float(u & 0xFFFF0000u), lab == c + 1, alpha, gamma, &d1);
    if (keep && !keep[p]) d0 = d1 = 0.f;
    __nv_bfloat162 h = __floats2bfloat162_rn(d0 * coef, d1 * coef);
    *reinterpret_cast<__nv_bfloat162*>(dlogits + p * ld + c) = h;
  }
}

// ------------------------------------------------------------------------------------ positives
struct PosOut {      // per-location forward terms
  float bce, giou_w, nll, l1, sel;
};

__device__ __forceinline__ float min_grad(float a, float b) { return a < b ? 1.f : (a == b ? 0.5f : 0.f); }
__device__ __forceinline__ float max_grad(float a, float b) { return a > b ? 1.f : (a == b ? 0.5f : 0.f); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// mode 0: labeled (bce + giou*ctr + nll*iou); 1: pseudo-cls set (bce); 2: pseudo-reg set (ts-better masked L1)
// When `grad` is non-null the per-row gradient wrt the 73 raw head outputs is written (already scaled by
// the coefficients c_*), and dscale receives sum_i r_i * dz_i.
template <bool BWD>
__device__ __forceinline__ PosOut pos_terms(const bf16* __restrict__ row, float scale, const float (&t)[4],
                                            const float (&bvar)[4], int mode, int kl_mode, float ts_better, float ts_cert,
                                            float c_bce, float c_giou, float c_nll, float c_l1, float* grow,
                                            float* dscale) {
  PosOut o = {0.f, 0.f, 0.f, 0.f, 0.f};
  const float ctr_logit = __bfloat162float(row[72]);
  const float ct = ctr_target(t);
  if (mode == 0 || mode == 1) {
    // BCE with logits: max(x,0) - x*t + log1p(exp(-|x|))
    o.bce = fmaxf(ctr_logit, 0.f) - ctr_logit * ct + log1pf(expf(-fabsf(ctr_logit)));
    if (BWD) grow[72] = c_bce * (sigmoidf_(ctr_logit) - ct);
  }
  if (mode == 1) return o;
  // Integral over the 4 x 17 distribution logits (z = scale * r)
  float pred[4], dpred[4] = {0.f, 0.f, 0.f, 0.f};
  float std_[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float mx = -3.0e38f;
    for (int i = 0; i < 17; ++i) mx = fmaxf(mx, scale * __bfloat162float(row[k * 17 + i]));
    float den = 0.f, num = 0.f;
    for (int i = 0; i < 17; ++i) {
      const float e = expf(scale * __bfloat162float(row[k * 17 + i]) - mx);
      den += e;
      num += e * (float)i;
    }
    pred[k] = num / den;
    std_[k] = __bfloat162float(row[68 + k]);
  }
  if (mode == 0) {
    // ltrb IoU / GIoU terms
    const float ta = (t[0] + t[2]) * (t[1] + t[3]);
    const float pa = (pred[0] + pred[2]) * (pred[1] + pred[3]);
    const float wi = fminf(pred[0], t[0]) + fminf(pred[2], t[2]);
    const float hi = fminf(pred[3], t[3]) + fminf(pred[1], t[1]);
    const float gw = fmaxf(pred[0], t[0]) + fmaxf(pred[2], t[2]);
    const float gh = fmaxf(pred[3], t[3]) + fmaxf(pred[1], t[1]);
    const float inter = wi * hi, uni = ta + pa - inter, ac = gw * gh;
    const float iou = (inter + 1.f) / (uni + 1.f);
    const float giou = iou - (ac - uni) / ac;
    o.giou_w = (1.f - giou) * ct;
    float nll = 0.f, dmu[4], dsd[4], wrow = iou;
    if (kl_mode == 0) {
      // NLL: sum_k (t-mu)^2 / (2 s^2) + 0.5 log s^2, + 2 log(2 pi), times IoU (detached)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float sg = sigmoidf_(std_[k]);
        const float s2 = sg * sg;
        const float d = t[k] - pred[k];
        nll += d * d / (2.f * s2) + 0.5f * logf(s2);
        dmu[k] = -d / s2;
        dsd[k] = (1.f - d * d / s2) * (1.f - sg);
      }
      nll += 2.f * logf(2.f * 3.14159265358979323846f);
    } else {
      // KLLoss (layers/kl_loss.py:17-66, beta = 1): exp(-s) * smooth_l1(mu - t) + s / 2 on the RAW uncertainty output, summed over
      // the four sides; the centerness target weights the row for the weight_ctr_* reductions (kl_mode 3, 4)
      wrow = kl_mode >= 3 ? ct : 1.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float d = pred[k] - t[k], n = fabsf(d);
        const float sl1 = n < 1.f ? 0.5f * n * n : n - 0.5f;
        const float e = expf(-std_[k]);
        nll += e * sl1 + 0.5f * std_[k];
        dmu[k] = e * (n < 1.f ? d : (d > 0.f ? 1.f : -1.f));
        dsd[k] = 0.5f - e * sl1;
      }
    }
    o.nll = nll * wrow;
    if (BWD) {
      // d(1-giou)/dpred
      const float dwi[4] = {min_grad(pred[0], t[0]), 0.f, min_grad(pred[2], t[2]), 0.f};
      const float dhi[4] = {0.f, min_grad(pred[1], t[1]), 0.f, min_grad(pred[3], t[3])};
      const float dgw[4] = {max_grad(pred[0], t[0]), 0.f, max_grad(pred[2], t[2]), 0.f};
      const float dgh[4] = {0.f, max_grad(pred[1], t[1]), 0.f, max_grad(pred[3], t[3])};
      const float dpa[4] = {pred[1] + pred[3], pred[0] + pred[2], pred[1] + pred[3], pred[0] + pred[2]};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float dinter = dwi[k] * hi + wi * dhi[k];
        const float duni = dpa[k] - dinter;
        const float dac = dgw[k] * gh + gw * dgh[k];
        const float diou = (dinter * (uni + 1.f) - (inter + 1.f) * duni) / ((uni + 1.f) * (uni + 1.f));
        const float dgiou = diou + (duni * ac - uni * dac) / (ac * ac);
        dpred[k] = -c_giou * ct * dgiou + c_nll * wrow * dmu[k];
        grow[68 + k] = c_nll * wrow * dsd[k];
      }
    }
  } else {  // mode 2
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float cs = 1.f - sigmoidf_(std_[k]);
      const float ctch = 1.f - sigmoidf_(bvar[k]);
      const bool sel = (ctch > ts_cert) && (ctch > cs + ts_better);
      if (sel) {
        const float d = pred[k] - t[k];
        o.l1 += fabsf(d);
        o.sel += 1.f;
        if (BWD) dpred[k] = c_l1 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
      }
    }
  }
  if (BWD) {
    float ds = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (dpred[k] == 0.f) continue;   // rows start zeroed
      float mx = -3.0e38f;
      for (int i = 0; i < 17; ++i) mx = fmaxf(mx, scale * __bfloat162float(row[k * 17 + i]));
      float den = 0.f;
      for (int i = 0; i < 17; ++i) den += expf(scale * __bfloat162float(row[k * 17 + i]) - mx);
      for (int i = 0; i < 17; ++i) {
        const float r = __bfloat162float(row[k * 17 + i]);
        const float pi = expf(scale * r - mx) / den;
        const float dz = dpred[k] * pi * ((float)i - pred[k]);
        grow[k * 17 + i] = scale * dz;
        ds += r * dz;
      }
    }
    *dscale = ds;
  }
  return o;
}

__global__ void __launch_bounds__(128)
pos_fwd_kernel(Levels lv, int N, const bf16* __restrict__ box_out, int ld, const float* __restrict__ scales,
               const long long* __restrict__ labels, const float* __restrict__ reg_t, const float* __restrict__ bvar,
               int num_classes, int mode, int kl_mode, float ts_better, float ts_cert, double* __restrict__ acc) {
  const long long P = (long long)lv.off[lv.num] * N;
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (p < P && labels[p] != num_classes) {
    int l, img, hw;
    locate(lv, N, p, l, img, hw);
    const float4 t4 = reinterpret_cast<const float4*>(reg_t)[p];
    const float4 b4 = reinterpret_cast<const float4*>(bvar)[p];
    const float t[4] = {t4.x, t4.y, t4.z, t4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
    const PosOut o = pos_terms<false>(box_out + p * ld, scales[l], t, bv, mode, kl_mode, ts_better, ts_cert, 0, 0, 0, 0,
                                      nullptr, nullptr);
    v[0] = o.bce; v[1] = o.giou_w; v[2] = o.nll; v[3] = o.l1; v[4] = o.sel; v[5] = 1.f;
  }
  __shared__ float red[4][6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
  }
  if ((threadIdx.x & 31) == 0)
    for (int j = 0; j < 6; ++j) red[threadIdx.x >> 5][j] = v[j];
  __syncthreads();
  if (threadIdx.x == 0) {
    float s[6];
    for (int j = 0; j < 6; ++j) s[j] = red[0][j] + red[1][j] + red[2][j] + red[3][j];
    if (s[5] != 0.f)
      for (int j = 0; j < 6; ++j) atomicAdd(acc + 1 + j, (double)s[j]);   // acc[1..6]
  }
}

// losses[0..3] = cls, loc, ctr, teacher_better_student  (mode semantics as above)
// divisor of the summed uncertainty term: NLLoss = mean over the positives; KLLoss by LOC_FUN_ALL (kl_mode 1 mean over
// 4 * positives, 2 sum, 3 weight_ctr_sum, 4 weight_ctr_mean = / loss_denorm)
__device__ __forceinline__ double kl_divisor(int kl_mode, double npos, float denorm) {
  return kl_mode == 0 ? npos : (kl_mode == 1 ? 4.0 * npos : (kl_mode == 4 ? (double)denorm : 1.0));
}

__global__ void finalize_kernel(const double* __restrict__ acc, const float* __restrict__ norm, float world, int mode,
                                int kl_mode, float kl_w, float* __restrict__ losses) {
  const float num_pos_avg = fmaxf(norm[0] / world, 1.0f);
  const float denorm = fmaxf(norm[1] / world, 1e-6f);
  const double npos = acc[6];
  float cls = (float)(acc[0] / num_pos_avg), loc = 0.f, ctr = 0.f, tbs = 0.f;
  if (npos > 0.0) {
    ctr = (float)(acc[1] / num_pos_avg);
    if (mode == 0) loc = (float)(kl_w * kl_w * (acc[3] / kl_divisor(kl_mode, npos, denorm)) + acc[2] / denorm);
    if (mode == 2) { loc = acc[5] > 0.0 ? (float)(acc[4] / acc[5]) : 0.f; tbs = (float)acc[5]; }
  }
  if (mode == 0 && npos == 0.0) cls = 0.f;    // fcos_outputs.py:430-434
  losses[0] = cls; losses[1] = loc; losses[2] = ctr; losses[3] = tbs;
}

__global__ void __launch_bounds__(128)
pos_bwd_kernel(Levels lv, int N, const bf16* __restrict__ box_out, int ld, const float* __restrict__ scales,
               const long long* __restrict__ labels, const float* __restrict__ reg_t, const float* __restrict__ bvar,
               int num_classes, int mode, int kl_mode, float ts_better, float ts_cert, float kl_w, const double* __restrict__ acc,
               const float* __restrict__ norm, float world, const float* __restrict__ gout /*[cls, loc, ctr]*/,
               bf16* __restrict__ dbox, float* __restrict__ dscales, int accumulate) {
  const long long P = (long long)lv.off[lv.num] * N;
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= P) return;
  uint4* drow = reinterpret_cast<uint4*>(dbox + p * ld);
  if (labels[p] == num_classes) {
    if (!accumulate) {
#pragma unroll
      for (int j = 0; j < 10; ++j) drow[j] = make_uint4(0, 0, 0, 0);
    }
    return;
  }
  const float num_pos_avg = fmaxf(norm[0] / world, 1.0f);
  const float denorm = fmaxf(norm[1] / world, 1e-6f);
  const float npos = (float)acc[6];
  const float c_bce = gout[2] / num_pos_avg;
  const float c_giou = gout[1] / denorm;
  const float c_nll = gout[1] * kl_w * kl_w / (float)kl_divisor(kl_mode, (double)npos, denorm);
  const float c_l1 = acc[5] > 0.0 ? gout[1] / (float)acc[5] : 0.f;
  int l, img, hw;
  locate(lv, N, p, l, img, hw);
  const float4 t4 = reinterpret_cast<const float4*>(reg_t)[p];
  const float4 b4 = reinterpret_cast<const float4*>(bvar)[p];
  const float t[4] = {t4.x, t4.y, t4.z, t4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
  float grow[80];
#pragma unroll
  for (int j = 0; j < 80; ++j) grow[j] = 0.f;
  float ds = 0.f;
  pos_terms<true>(box_out + p * ld, scales[l], t, bv, mode, kl_mode, ts_better, ts_cert, c_bce, c_giou, c_nll, c_l1, grow, &ds);
  if (accumulate) {
    const bf16* old = dbox + p * ld;
#pragma unroll
    for (int j = 0; j < 80; ++j) grow[j] += __bfloat162float(old[j]);
  }
#pragma unroll
  for (int j = 0; j < 10; ++j) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      __nv_bfloat162 h = __floats2bfloat162_rn(grow[8 * j + 2 * q], grow[8 * j + 2 * q + 1]);
      w[q] = *reinterpret_cast<uint32_t*>(&h);
    }
    drow[j] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (ds != 0.f && dscales) atomicAdd(dscales + l, ds);
}

int fill_levels(Levels& lv, int num_levels, const int* hw, const int* strides, const float* ranges) {
  if (num_levels < 1 || num_levels > MAXL) return -1;
  lv.num = num_levels;
  lv.off[0] = 0;
  for (int i = 0; i < num_levels; ++i) {
    lv.H[i] = hw[2 * i]; lv.W[i] = hw[2 * i + 1]; lv.stride[i] = strides[i];
    lv.lo[i] = ranges ? ranges[2 * i] : 0.f; lv.hi[i] = ranges ? ranges[2 * i + 1] : 0.f;
    lv.off[i + 1] = lv.off[i] + hw[2 * i] * hw[2 * i + 1];
  }
  return 0;
}
}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

// hw / strides / ranges are HOST arrays (level geometry is a launch parameter, not data).
extern "C" int ut2_fcos_assign_targets(int num_levels, const int* hw, const int* strides, const float* ranges, int N,
                                       int G, const float* boxes, const long long* classes, const int* counts,
                                       const float* bvar, int num_classes, float center_radius, int ignore_near,
                                       long long* labels, long long* tinds, float* reg_t, float* bv_out,
                                       unsigned char* keep, float* norm, void* stream) {
  Levels lv;
  if (fill_levels(lv, num_levels, hw, strides, ranges)) return ut2_fail(-2, "assign_targets: bad level count");
  const long long P = (long long)lv.off[lv.num] * N;
  if (P <= 0) return ut2_fail(-2, "assign_targets: empty");
  cudaMemsetAsync(norm, 0, 2 * sizeof(float), STREAM);
  assign_targets_kernel<<<ut2_ceil_div(P, 256), 256, 0, STREAM>>>(lv, N, G, boxes, classes, counts, bvar, num_classes,
                                                                  center_radius, ignore_near, labels, tinds, reg_t, bv_out, keep, norm);
  return ut2_check_launch("fcos_assign_targets");
}

// acc: double[8] zeroed here; mode 0/1 also run the focal sum over cls_out. losses: float[4].
extern "C" int ut2_fcos_loss_fwd(int num_levels, const int* hw, const int* strides, int N, const void* cls_out,
                                 const void* box_out, int ld, const float* scales, const long long* labels,
                                 const unsigned char* keep, const float* reg_t, const float* bvar, int num_classes,
                                 int mode, float alpha, float gamma, float kl_w, float ts_better, float ts_cert,
                                 const float* norm, float world, int kl_mode, double* acc, float* losses, void* stream) {
  if (kl_mode < 0 || kl_mode > 4) return ut2_fail(-2, "fcos_loss_fwd: kl_mode 0 (nlloss) or 1..4 (klloss: mean, sum, weight_ctr_sum, weight_ctr_mean)");
  Levels lv;
  if (fill_levels(lv, num_levels, hw, strides, nullptr)) return ut2_fail(-2, "fcos_loss_fwd: bad level count");
  const long long P = (long long)lv.off[lv.num] * N;
  cudaMemsetAsync(acc, 0, 8 * sizeof(double), STREAM);
  if (mode != 2) {
    const long long total = P * (num_classes / 2);
    long long g = (total + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    focal_fwd_kernel<<<(int)g, 256, 0, STREAM>>>(static_cast<const bf16*>(cls_out), ld, num_classes, labels,
                                                 mode == 0 ? keep : nullptr, P, alpha, gamma, acc);
  }
  pos_fwd_kernel<<<ut2_ceil_div(P, 128), 128, 0, STREAM>>>(lv, N, static_cast<const bf16*>(box_out), ld, scales, labels,
                                                           reg_t, bvar, num_classes, mode, kl_mode, ts_better, ts_cert, acc);
  finalize_kernel<<<1, 1, 0, STREAM>>>(acc, norm, world, mode, kl_mode, kl_w, losses);
  return ut2_check_launch("fcos_loss_fwd");
}

// gout: device float[3] = d(total)/d(cls, loc, ctr). dcls / dbox are fully written (zeros off the positives);
// dcls may be null for mode 2, dbox may be null never. dscales (float[num_levels]) is accumulated.
// accumulate != 0: dbox rows of the positives are added to (used to merge the pseudo cls-set and reg-set grads).
extern "C" int ut2_fcos_loss_bwd(int num_levels, const int* hw, const int* strides, int N, const void* cls_out,
                                 const void* box_out, int ld, const float* scales, const long long* labels,
                                 const unsigned char* keep, const float* reg_t, const float* bvar, int num_classes,
                                 int mode, float alpha, float gamma, float kl_w, float ts_better, float ts_cert,
                                 const float* norm, float world, int kl_mode, const double* acc, const float* gout, void* dcls,
                                 void* dbox, float* dscales, int accumulate, void* stream) {
  if (kl_mode < 0 || kl_mode > 4) return ut2_fail(-2, "fcos_loss_bwd: kl_mode 0..4");
  Levels lv;
  if (fill_levels(lv, num_levels, hw, strides, nullptr)) return ut2_fail(-2, "fcos_loss_bwd: bad level count");
  const long long P = (long long)lv.off[lv.num] * N;
  if (mode != 2 && dcls) {
    const long long total = P * (num_classes / 2);
    long long g = (total + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    focal_bwd_kernel<<<(int)g, 256, 0, STREAM>>>(static_cast<const bf16*>(cls_out), ld, num_classes, labels,
                                                 mode == 0 ? keep : nullptr, P, alpha, gamma, norm, world, gout, acc,
                                                 mode == 0, static_cast<bf16*>(dcls));
  }
  pos_bwd_kernel<<<ut2_ceil_div(P, 128), 128, 0, STREAM>>>(lv, N, static_cast<const bf16*>(box_out), ld, scales, labels,
                                                           reg_t, bvar, num_classes, mode, kl_mode, ts_better, ts_cert, kl_w, acc,
                                                           norm, world, gout, static_cast<bf16*>(dbox), dscales,
                                                           accumulate);
  return ut2_check_launch("fcos_loss_bwd");
}
