// ROI heads of the Faster R-CNN half of Unbiased Teacher v2: proposal labelling + sampling, ROIAlign (forward and
// backward), the box-predictor losses (forward and backward) and the detection candidates of the teacher.
//
// Reference (paths under /root/reference/ubteacher; [D2] = Detectron2 v0.6, [tv] = torchvision):
//   modeling/roi_heads/roi_heads.py:138-270  label_and_sample_proposals[_pseudo]  ([D2] add_ground_truth_to_proposals,
//                                            pairwise_iou, Matcher(.5), _sample_proposals / subsample_labels)
//   roi_heads.py:110-136                     _forward_box -> [D2] ROIPooler(ROIAlignV2 7x7, sampling_ratio 0) -> [tv] roi_align
//   modeling/roi_heads/fast_rcnn.py:834-1084 FastRCNNFocaltLossBoundaryVarOutputLayers.losses (FocalLoss :1405, nl_loss
//                                            :1228, matched_boxlist_iou :20, Box2BoxXYXYTransform box_regression.py:12-129)
//   fast_rcnn.py:1086-1125                   inference -> [D2] fast_rcnn_inference (softmax, decode, clip, > 0.05, NMS .5, 100)
//
// Layouts: features NHWC bf16 per level; pooled ROI features [R, 7, 7, C] bf16 (== the [R, 12544] GEMM operand of fc1,
// k = (ph*7 + pw)*C + c); fused predictor output [R, 96] bf16: 0..80 class scores | 81..84 deltas (l, r, d, u) |
// 85..88 delta std | pad. ROIs are fixed-capacity [N, Rcap] rows + a per-image count; rows >= count are inert.
#include <stdlib.h>
#include "ut2_internal.h"
#include <cuda_bf16.h>
#include <stdint.h>

namespace {
typedef __nv_bfloat16 bf16;
constexpr int GMAX = 128;
constexpr int SAMPLE_CAP = 2048;     // proposals + appended ground truth per image
constexpr int PLD = 96;              // columns of the fused box-predictor output
constexpr int NCLS = 80;

__device__ __forceinline__ float iou_pair(const float4 g, const float4 a) {
  const float w = fmaxf(__fsub_rn(fminf(g.z, a.z), fmaxf(g.x, a.x)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(g.w, a.w), fmaxf(g.y, a.y)), 0.f);
  const float inter = __fmul_rn(w, h);
  if (!(inter > 0.f)) return 0.f;
  const float ag = __fmul_rn(__fsub_rn(g.z, g.x), __fsub_rn(g.w, g.y));
  const float aa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(ag, aa), inter));
}

__device__ __forceinline__ uint32_t hash_key(uint32_t seed, uint32_t img, uint32_t idx) {
  uint32_t x = seed ^ (img * 0x9E3779B9u) ^ (idx * 0x85EBCA6Bu);
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  x += idx * 0xC2B2AE35u; x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12;
  return x;
}

// ------------------------------------------------------------------------------------ 1. label + sample proposals
// One CTA per image. Candidates = first prop_cnt proposals followed by the gt_cnt ground-truth boxes.
__global__ void __launch_bounds__(1024)
roi_sample_kernel(int Pcap, int G, int Rcap, const float* __restrict__ prop_boxes, const int* __restrict__ prop_cnt,
                  const float* __restrict__ gt_boxes, const long long* __restrict__ gt_classes, const int* __restrict__ gt_cnt,
                  const float* __restrict__ gt_scores, const float* __restrict__ gt_std, const uint32_t* __restrict__ keys,
                  int key_ld, uint32_t seed, const uint32_t* __restrict__ seed_dev, int max_fg, float iou_thr, int num_classes,
                  int append_gt, float* __restrict__ roi_box, long long* __restrict__ roi_cls, float* __restrict__ roi_gtbox,
                  float* __restrict__ roi_conf, float* __restrict__ roi_std, int* __restrict__ roi_src, int* __restrict__ roi_cnt) {
  __shared__ float4 sg[GMAX];
  __shared__ unsigned long long skey[SAMPLE_CAP];
  __shared__ short smatch[SAMPLE_CAP];
  __shared__ unsigned char sfg[SAMPLE_CAP];
  __shared__ int s_nfg, s_nbg;
  const int img = blockIdx.x;
  if (seed_dev) seed += seed_dev[0] * 0x9E3779B1u;     // device-resident draw counter (CUDA-graph replay)
  const int ng = min(gt_cnt[img], G);
  const int np = min(prop_cnt[img], Pcap);
  const int M = min(np + (append_gt ? ng : 0), SAMPLE_CAP);
  for (int i = threadIdx.x; i < ng; i += blockDim.x) sg[i] = reinterpret_cast<const float4*>(gt_boxes)[(size_t)img * G + i];
  if (threadIdx.x == 0) { s_nfg = 0; s_nbg = 0; }
  __syncthreads();
  auto box_of = [&](int i) -> float4 {
    return i < np ? reinterpret_cast<const float4*>(prop_boxes)[(size_t)img * Pcap + i] : sg[i - np];
  };
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const float4 b = box_of(i);
    float best = -1.f;
    int bi = 0;
    for (int g = 0; g < ng; ++g) {
      const float v = iou_pair(sg[g], b);
      if (v > best) { best = v; bi = g; }
    }
    const bool fg = ng > 0 && best >= iou_thr;          // Matcher(.5): [0.5, inf) -> 1, no low-quality matches
    smatch[i] = (short)bi;
    sfg[i] = fg;
    const uint32_t k = keys ? keys[(size_t)img * key_ld + i] : hash_key(seed, (uint32_t)img, (uint32_t)i);
    skey[i] = ((unsigned long long)k << 32) | (uint32_t)i;
    atomicAdd(fg ? &s_nfg : &s_nbg, 1);
  }
  __syncthreads();
  const int n_fg = min(s_nfg, max_fg);
  const int n_bg = min(s_nbg, Rcap - n_fg);
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const bool fg = sfg[i];
    const unsigned long long k = skey[i];
    int rank = 0;
    for (int j = 0; j < M; ++j) rank += (sfg[j] == (unsigned char)fg) && (skey[j] < k);
    const int want = fg ? n_fg : n_bg;
    if (rank >= want) continue;
    const size_t o = (size_t)img * Rcap + (fg ? rank : n_fg + rank);
    const int gi = smatch[i];
    reinterpret_cast<float4*>(roi_box)[o] = box_of(i);
    roi_cls[o] = fg ? gt_classes[(size_t)img * G + gi] : (long long)num_classes;
    reinterpret_cast<float4*>(roi_gtbox)[o] = ng > 0 ? sg[gi] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (roi_conf) roi_conf[o] = (ng > 0 && gt_scores) ? gt_scores[(size_t)img * G + gi] : 0.f;
    if (roi_std)
      reinterpret_cast<float4*>(roi_std)[o] = (ng > 0 && gt_std) ? reinterpret_cast<const float4*>(gt_std)[(size_t)img * G + gi]
                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    roi_src[o] = i;
  }
  for (int r = n_fg + n_bg + threadIdx.x; r < Rcap; r += blockDim.x) {      // inert tail
    const size_t o = (size_t)img * Rcap + r;
    reinterpret_cast<float4*>(roi_box)[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    roi_cls[o] = -1;
    reinterpret_cast<float4*>(roi_gtbox)[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (roi_conf) roi_conf[o] = 0.f;
    if (roi_std) reinterpret_cast<float4*>(roi_std)[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    roi_src[o] = -1;
  }
  if (threadIdx.x == 0) roi_cnt[img] = n_fg + n_bg;
}

// ------------------------------------------------------------------------------------ 2. ROIAlign
struct RoiLevels {
  int num;
  const bf16* feat[4];
  float* dfeat[4];
  int H[4], W[4];
  float scale[4];
};

// [D2] assign_boxes_to_levels: floor(4 + log2(sqrt(area) / 224 + 1e-8)) clamped to [2, 5], minus 2
__device__ __forceinline__ int roi_level(const float4 b, int num) {
  const float size = sqrtf((b.z - b.x) * (b.w - b.y));
  float lv = floorf(4.f + log2f(size / 224.f + 1e-8f));
  lv = fminf(fmaxf(lv, 2.f), 5.f);
  int l = (int)lv - 2;
  return l < num ? l : num - 1;
}

// One CTA per ROI, warp `ph` handles output row ph, lane handles 8 channels per 256-channel slice.
// [tv] roi_align, aligned=True, sampling_ratio=0: grid = ceil(roi_size / 7) samples per bin and axis.
template <bool BWD>
__global__ void __launch_bounds__(224)
roi_align_kernel(RoiLevels lv, int Rcap, int C, const float* __restrict__ rois, const int* __restrict__ roi_cnt,
                 bf16* __restrict__ out, const bf16* __restrict__ dout) {
  const int r = blockIdx.x;
  const int img = r / Rcap, k = r - img * Rcap;
  const int ph = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool live = k < roi_cnt[img];
  bf16* orow = BWD ? nullptr : out + ((size_t)r * 49 + ph * 7) * C;
  if (!live) {
    if (!BWD)
      for (int i = lane; i < 7 * C / 8; i += 32) reinterpret_cast<uint4*>(orow)[i] = make_uint4(0, 0, 0, 0);
    return;
  }
  const float4 b = reinterpret_cast<const float4*>(rois)[r];
  const int l = roi_level(b, lv.num);
  const int H = lv.H[l], W = lv.W[l];
  const float sc = lv.scale[l];
  const float x1 = b.x * sc - 0.5f, y1 = b.y * sc - 0.5f;
  const float rw = b.z * sc - 0.5f - x1, rh = b.w * sc - 0.5f - y1;
  const float bw = rw / 7.f, bh = rh / 7.f;
  const int gh = (int)ceilf(rh / 7.f), gw = (int)ceilf(rw / 7.f);
  const float inv_cnt = 1.f / fmaxf((float)(gh * gw), 1.f);
  const bf16* f = lv.feat[l] + (size_t)img * H * W * C;
  float* df = BWD ? lv.dfeat[l] + (size_t)img * H * W * C : nullptr;
  for (int pw = 0; pw < 7; ++pw) {
    for (int c0 = lane * 8; c0 < C; c0 += 256) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      if (BWD) {
        const uint4 g = *reinterpret_cast<const uint4*>(dout + ((size_t)r * 49 + ph * 7 + pw) * C + c0);
        const uint32_t gw4[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[2 * i] = __uint_as_float(gw4[i] << 16) * inv_cnt;
          acc[2 * i + 1] = __uint_as_float(gw4[i] & 0xFFFF0000u) * inv_cnt;
        }
      }
      for (int iy = 0; iy < gh; ++iy) {
        float y = y1 + ph * bh + (iy + 0.5f) * bh / (float)gh;
        for (int ix = 0; ix < gw; ++ix) {
          float x = x1 + pw * bw + (ix + 0.5f) * bw / (float)gw;
          if (y < -1.f || y > (float)H || x < -1.f || x > (float)W) continue;
          float yy = fmaxf(y, 0.f), xx = fmaxf(x, 0.f);
          int yl = (int)yy, xl = (int)xx, yh, xh;
          if (yl >= H - 1) { yh = yl = H - 1; yy = (float)yl; } else yh = yl + 1;
          if (xl >= W - 1) { xh = xl = W - 1; xx = (float)xl; } else xh = xl + 1;
          const float ly = yy - yl, lx = xx - xl, hy = 1.f - ly, hx = 1.f - lx;
          const float w4[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
          const size_t o4[4] = {((size_t)yl * W + xl) * C + c0, ((size_t)yl * W + xh) * C + c0,
                                ((size_t)yh * W + xl) * C + c0, ((size_t)yh * W + xh) * C + c0};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (BWD) {
              float* d = df + o4[q];
              atomicAdd(reinterpret_cast<float4*>(d), make_float4(acc[0] * w4[q], acc[1] * w4[q], acc[2] * w4[q], acc[3] * w4[q]));
              atomicAdd(reinterpret_cast<float4*>(d + 4), make_float4(acc[4] * w4[q], acc[5] * w4[q], acc[6] * w4[q], acc[7] * w4[q]));
            } else {
              const uint4 v = __ldg(reinterpret_cast<const uint4*>(f + o4[q]));
              const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                acc[2 * i] += w4[q] * __uint_as_float(vw[i] << 16);
                acc[2 * i + 1] += w4[q] * __uint_as_float(vw[i] & 0xFFFF0000u);
              }
            }
          }
        }
      }
      if (!BWD) {
        uint32_t ow[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __nv_bfloat162 hh = __floats2bfloat162_rn(acc[2 * i] * inv_cnt, acc[2 * i + 1] * inv_cnt);
          ow[i] = *reinterpret_cast<uint32_t*>(&hh);
        }
        *reinterpret_cast<uint4*>(orow + (size_t)pw * C + c0) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
    }
  }
}

// Backward with the sample points of a bin merged per feature pixel. The bilinear weight of a sample is separable,
// w(sample, pixel) = wy(iy, py) * wx(ix, px), and so is the validity test, hence the total weight of pixel (py, px) in bin
// (ph, pw) is Wy[ph][py] * Wx[pw][px] with Wy / Wx the per-axis sums over the bin's samples: (gh + 1)(gw + 1) vector atomics per
// bin and lane instead of 4 gh gw (gh = gw = 2..4 for ROIs on their canonical level: 16..64 -> 9..25). Warp w builds the row table
// of output row w and the column table of output column w in shared memory (lane e = pixel offset e); ROIs whose bins span
// more than 32 pixel rows / columns (degenerate aspect ratios) take the per-sample path of roi_align_kernel<true>.
constexpr int RA_T = 32;
struct AxisSample { int lo, hi; float wl, wh; bool ok; };
__device__ __forceinline__ AxisSample axis_sample(float v, int L) {
  AxisSample r;
  r.ok = !(v < -1.f || v > (float)L);
  float vv = fmaxf(v, 0.f);
  r.lo = (int)vv;
  if (r.lo >= L - 1) { r.hi = r.lo = L - 1; vv = (float)r.lo; } else r.hi = r.lo + 1;
  r.wh = vv - r.lo;
  r.wl = 1.f - r.wh;
  return r;
}
// pixel range [lo, lo + n) and weights of bin `b` along one axis (start v1, bin size bs, g samples, L pixels)
__device__ __forceinline__ void axis_table(float v1, float bs, int g, int L, int b, int lane, float* tab, int* lo_out, int* n_out) {
  int lo = 1 << 30, hi = -1;
  for (int i = 0; i < g; ++i) {
    const AxisSample a = axis_sample(v1 + b * bs + (i + 0.5f) * bs / (float)g, L);
    if (a.ok) { lo = min(lo, a.lo); hi = max(hi, a.hi); }
  }
  const int n = hi >= lo ? hi - lo + 1 : 0;
  if (n > 0 && n <= RA_T && lane < n) {
    const int p = lo + lane;
    float w = 0.f;
    for (int i = 0; i < g; ++i) {
      const AxisSample a = axis_sample(v1 + b * bs + (i + 0.5f) * bs / (float)g, L);
      if (a.ok) w += (a.lo == p ? a.wl : 0.f) + (a.hi == p ? a.wh : 0.f);
    }
    tab[lane] = w;
  }
  if (lane == 0) { *lo_out = lo; *n_out = n; }
}

__global__ void __launch_bounds__(224)
roi_align_bwd_kernel(RoiLevels lv, int Rcap, int C, const float* __restrict__ rois, const int* __restrict__ roi_cnt,
                     const bf16* __restrict__ dout) {
  const int r = blockIdx.x;
  const int img = r / Rcap, k = r - img * Rcap;
  const int ph = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (k >= roi_cnt[img]) return;
  const float4 b = reinterpret_cast<const float4*>(rois)[r];
  const int l = roi_level(b, lv.num);
  const int H = lv.H[l], W = lv.W[l];
  const float sc = lv.scale[l];
  const float x1 = b.x * sc - 0.5f, y1 = b.y * sc - 0.5f;
  const float rw = b.z * sc - 0.5f - x1, rh = b.w * sc - 0.5f - y1;
  const float bw = rw / 7.f, bh = rh / 7.f;
  const int gh = (int)ceilf(rh / 7.f), gw = (int)ceilf(rw / 7.f);
  const float inv_cnt = 1.f / fmaxf((float)(gh * gw), 1.f);
  float* df = lv.dfeat[l] + (size_t)img * H * W * C;
  __shared__ float sWy[7][RA_T], sWx[7][RA_T];
  __shared__ int sYlo[7], sYn[7], sXlo[7], sXn[7];
  axis_table(y1, bh, gh, H, ph, lane, sWy[ph], &sYlo[ph], &sYn[ph]);
  axis_table(x1, bw, gw, W, ph, lane, sWx[ph], &sXlo[ph], &sXn[ph]);
  __syncthreads();
  bool merged = true;
#pragma unroll
  for (int i = 0; i < 7; ++i) merged = merged && sYn[i] <= RA_T && sXn[i] <= RA_T;
  const int ylo = sYlo[ph], yn = sYn[ph];
  for (int pw = 0; pw < 7; ++pw) {
    const int xlo = sXlo[pw], xn = sXn[pw];
    for (int c0 = lane * 8; c0 < C; c0 += 256) {
      float acc[8];
      const uint4 g = *reinterpret_cast<const uint4*>(dout + ((size_t)r * 49 + ph * 7 + pw) * C + c0);
      const uint32_t gw4[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] = __uint_as_float(gw4[i] << 16) * inv_cnt;
        acc[2 * i + 1] = __uint_as_float(gw4[i] & 0xFFFF0000u) * inv_cnt;
      }
      if (merged) {
        for (int iy = 0; iy < yn; ++iy) {
          const float wy = sWy[ph][iy];
          if (wy == 0.f) continue;
          float* drow = df + ((size_t)(ylo + iy) * W + xlo) * C + c0;
          for (int ix = 0; ix < xn; ++ix) {
            const float w = wy * sWx[pw][ix];
            if (w == 0.f) continue;
            float* d = drow + (size_t)ix * C;
            atomicAdd(reinterpret_cast<float4*>(d), make_float4(acc[0] * w, acc[1] * w, acc[2] * w, acc[3] * w));
            atomicAdd(reinterpret_cast<float4*>(d + 4), make_float4(acc[4] * w, acc[5] * w, acc[6] * w, acc[7] * w));
          }
        }
      } else {
        for (int iy = 0; iy < gh; ++iy) {
          const AxisSample ay = axis_sample(y1 + ph * bh + (iy + 0.5f) * bh / (float)gh, H);
          if (!ay.ok) continue;
          for (int ix = 0; ix < gw; ++ix) {
            const AxisSample ax = axis_sample(x1 + pw * bw + (ix + 0.5f) * bw / (float)gw, W);
            if (!ax.ok) continue;
            const float w4[4] = {ay.wl * ax.wl, ay.wl * ax.wh, ay.wh * ax.wl, ay.wh * ax.wh};
            const size_t o4[4] = {((size_t)ay.lo * W + ax.lo) * C + c0, ((size_t)ay.lo * W + ax.hi) * C + c0,
                                  ((size_t)ay.hi * W + ax.lo) * C + c0, ((size_t)ay.hi * W + ax.hi) * C + c0};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float* d = df + o4[q];
              atomicAdd(reinterpret_cast<float4*>(d), make_float4(acc[0] * w4[q], acc[1] * w4[q], acc[2] * w4[q], acc[3] * w4[q]));
              atomicAdd(reinterpret_cast<float4*>(d + 4), make_float4(acc[4] * w4[q], acc[5] * w4[q], acc[6] * w4[q], acc[7] * w4[q]));
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------ 3. box-predictor losses
// Box2BoxXYXYTransform (box_regression.py:36-129), weights (wx, wy) = (10, 10): deltas (l, r, d, u) act on (x1, x2, y1, y2).
struct XyxyT {
  float wx, wy, clamp;
};

// One warp per ROI row. mode 0: supervised ('nlloss': L1 + 0.05 * NLL * IoU), mode 1: pseudo ('tsbetter' masked L1).
// acc: double[2] = {sum focal, sum box}; rtot = number of live rows (normaliser R).
template <bool BWD>
__global__ void __launch_bounds__(256)
fastrcnn_loss_kernel(int Rtot, int Rcap, const bf16* __restrict__ pred, const float* __restrict__ rois,
                     const long long* __restrict__ gt_cls, const float* __restrict__ gt_box, const float* __restrict__ gt_std,
                     const int* __restrict__ roi_cnt, int N, int mode, XyxyT T, float gamma, float nll_w, float ts_better,
                     float t_cert, const float* __restrict__ gout, double* __restrict__ acc, bf16* __restrict__ dpred) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  float s_cls = 0.f, s_box = 0.f;
  if (row < Rtot) {
    const int img = row / Rcap, k = row - img * Rcap;
    const bool live = k < roi_cnt[img];
    float inv_R = 0.f;
    if (BWD) {
      int R = 0;
      for (int i = 0; i < N; ++i) R += roi_cnt[i];
      inv_R = 1.f / fmaxf((float)R, 1.f);
    }
    const bf16* p = pred + (size_t)row * PLD;
    const long long t = live ? gt_cls[row] : -1;
    float x[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) x[i] = (lane + 32 * i) <= NCLS ? __bfloat162float(p[lane + 32 * i]) : -INFINITY;
    float mx = fmaxf(fmaxf(x[0], x[1]), x[2]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) se += (lane + 32 * i) <= NCLS ? expf(x[i] - mx) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    const float lse = mx + logf(se);
    float d[3] = {0.f, 0.f, 0.f};
    if (live && t >= 0) {
      const float xt = __bfloat162float(p[t]);
      const float ce = lse - xt;
      const float pt = expf(-ce);
      const float om = fmaxf(1.f - pt, 0.f);
      if (BWD) {
        // d/dCE [(1-p)^g * CE] = (1-p)^g + g (1-p)^(g-1) p CE
        const float dce = powf(om, gamma) + (om > 0.f ? gamma * powf(om, gamma - 1.f) * pt * ce : 0.f);
        const float g = gout[0] * inv_R * dce;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int c = lane + 32 * i;
          if (c <= NCLS) d[i] = g * (expf(x[i] - lse) - (c == (int)t ? 1.f : 0.f));
        }
      } else if (lane == 0) {
        s_cls = powf(om, gamma) * ce;
      }
    }
    float dbox[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dbox[i] = 0.f;
    if (live && t >= 0 && t < NCLS && lane == 0) {
      const float4 pb = reinterpret_cast<const float4*>(rois)[row];
      const float4 gb = reinterpret_cast<const float4*>(gt_box)[row];
      float mu[4], sd[4], tg[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { mu[i] = __bfloat162float(p[81 + i]); sd[i] = __bfloat162float(p[85 + i]); }
      const float sw = pb.z - pb.x + 1.f, sh = pb.w - pb.y + 1.f;           // get_deltas: size + 1
      tg[0] = T.wx * (gb.x - pb.x) / sw; tg[1] = T.wx * (gb.z - pb.z) / sw;
      tg[2] = T.wy * (gb.y - pb.y) / sh; tg[3] = T.wy * (gb.w - pb.w) / sh;
      const float gs = BWD ? gout[1] * inv_R : 0.f;
      if (mode == 1) {
        const float4 ts = reinterpret_cast<const float4*>(gt_std)[row];
        const float tstd[4] = {ts.x, ts.y, ts.z, ts.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float ct = 1.f - 1.f / (1.f + expf(-tstd[i]));
          const float cs = 1.f - 1.f / (1.f + expf(-sd[i]));
          if (ct > cs + ts_better && ct > t_cert) {
            const float e = mu[i] - tg[i];
            if (BWD) dbox[i] = e > 0.f ? gs : (e < 0.f ? -gs : 0.f);
            else s_box += fabsf(e);
          }
        }
      } else {
        const float w = pb.z - pb.x, h = pb.w - pb.y;                        // apply_deltas: size without + 1
        const float wgt[4] = {T.wx, T.wx, T.wy, T.wy};
        const float size[4] = {w, w, h, h};
        const float base[4] = {pb.x, pb.z, pb.y, pb.w};
        float q[4];
        bool pass[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float v = mu[i] / wgt[i];
          pass[i] = v >= -T.clamp && v <= T.clamp;
          q[i] = fminf(fmaxf(v, -T.clamp), T.clamp) * size[i] + base[i];      // q = (x1, x2, y1, y2)
        }
        const float px1 = q[0], px2 = q[1], py1 = q[2], py2 = q[3];
        const float a1 = (gb.z - gb.x) * (gb.w - gb.y), a2 = (px2 - px1) * (py2 - py1);
        const float ltx = fmaxf(gb.x, px1), lty = fmaxf(gb.y, py1), rbx = fminf(gb.z, px2), rby = fminf(gb.w, py2);
        const float iw = fmaxf(rbx - ltx, 0.f), ih = fmaxf(rby - lty, 0.f);
        const float inter = iw * ih, U = a1 + a2 - inter;
        const float iou = inter / U;
        float S = 0.f, sig[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          sig[i] = 1.f / (1.f + expf(-sd[i]));
          const float e = tg[i] - mu[i];
          S += e * e / (2.f * sig[i] * sig[i]) + 0.5f * logf(sig[i] * sig[i]);
        }
        S += 2.f * logf(2.f * 3.14159265358979323846f);
        if (!BWD) {
          float l1 = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) l1 += fabsf(mu[i] - tg[i]);
          s_box = l1 + nll_w * S * iou;
        } else {
          // d iou / d (x1, x2, y1, y2) of the predicted box (the reference does not detach the IoU weight)
          const float live_w = (rbx - ltx) > 0.f ? 1.f : 0.f, live_h = (rby - lty) > 0.f ? 1.f : 0.f;
          const float di[4] = {px1 > gb.x ? -live_w * ih : 0.f, px2 < gb.z ? live_w * ih : 0.f,
                               py1 > gb.y ? -live_h * iw : 0.f, py2 < gb.w ? live_h * iw : 0.f};
          const float da[4] = {-(py2 - py1), (py2 - py1), -(px2 - px1), (px2 - px1)};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float diou = (di[i] * U - inter * (da[i] - di[i])) / (U * U);
            const float dq = pass[i] ? size[i] / wgt[i] : 0.f;
            const float e = mu[i] - tg[i];
            const float l1 = e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f);
            dbox[i] = gs * (l1 + nll_w * (iou * e / (sig[i] * sig[i]) + S * diou * dq));
            dbox[4 + i] = gs * nll_w * iou * (-(e * e) / (sig[i] * sig[i] * sig[i]) + 1.f / sig[i]) * sig[i] * (1.f - sig[i]);
          }
        }
      }
    }
    if (BWD) {
      bf16* o = dpred + (size_t)row * PLD;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = lane + 32 * i;
        if (c <= NCLS) o[c] = __float2bfloat16(d[i]);
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[81 + i] = __float2bfloat16(dbox[i]);
#pragma unroll
        for (int i = 89; i < PLD; ++i) o[i] = __float2bfloat16(0.f);
      }
    }
  }
  if (!BWD) {
    __shared__ float red[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s_cls += __shfl_xor_sync(0xffffffffu, s_cls, o);
      s_box += __shfl_xor_sync(0xffffffffu, s_box, o);
    }
    if (lane == 0) { red[0][threadIdx.x >> 5] = s_cls; red[1][threadIdx.x >> 5] = s_box; }
    __syncthreads();
    if (threadIdx.x < 2) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
      if (s != 0.f) atomicAdd(&acc[threadIdx.x], (double)s);
    }
  }
}

__global__ void fastrcnn_loss_finalize(const double* __restrict__ acc, const int* __restrict__ roi_cnt, int N,
                                       float* __restrict__ losses) {
  if (threadIdx.x < 2) {
    int R = 0;
    for (int i = 0; i < N; ++i) R += roi_cnt[i];
    losses[threadIdx.x] = R > 0 ? (float)(acc[threadIdx.x] / (double)R) : 0.f;
  }
}

// ------------------------------------------------------------------------------------ 4. detection candidates
// One warp per ROI row: softmax over 81 scores, class-agnostic decode, clip; every class with p > thr becomes a
// candidate (box, p, class, canon = row * 80 + class) in the image's list (capacity Ccap, overflow counted).
__global__ void __launch_bounds__(256)
fastrcnn_candidates_kernel(int Rtot, int Rcap, const bf16* __restrict__ pred, const float* __restrict__ rois,
                           const int* __restrict__ roi_cnt, const float* __restrict__ image_hw, XyxyT T, float thr, int Ccap,
                           float* __restrict__ cand_box, float* __restrict__ cand_score, int* __restrict__ cand_cls,
                           int* __restrict__ cand_canon, int* __restrict__ cand_cnt, int* __restrict__ overflow) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= Rtot) return;
  const int img = row / Rcap, k = row - img * Rcap;
  if (k >= roi_cnt[img]) return;
  const bf16* p = pred + (size_t)row * PLD;
  float x[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) x[i] = (lane + 32 * i) <= NCLS ? __bfloat162float(p[lane + 32 * i]) : -INFINITY;
  float mx = fmaxf(fmaxf(x[0], x[1]), x[2]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float e[3], se = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) { e[i] = (lane + 32 * i) <= NCLS ? expf(x[i] - mx) : 0.f; se += e[i]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
  const float4 pb = reinterpret_cast<const float4*>(rois)[row];
  const float w = pb.z - pb.x, h = pb.w - pb.y;
  const float ih = image_hw[img * 2], iw = image_hw[img * 2 + 1];
  auto dec = [&](int i, float wt, float size, float base) {
    const float v = fminf(fmaxf(__bfloat162float(p[81 + i]) / wt, -T.clamp), T.clamp);
    return v * size + base;
  };
  float x1 = dec(0, T.wx, w, pb.x), x2 = dec(1, T.wx, w, pb.z), y1 = dec(2, T.wy, h, pb.y), y2 = dec(3, T.wy, h, pb.w);
  const bool finite = isfinite(x1) && isfinite(x2) && isfinite(y1) && isfinite(y2) && isfinite(se) && isfinite(mx);
  x1 = fminf(fmaxf(x1, 0.f), iw); x2 = fminf(fmaxf(x2, 0.f), iw);
  y1 = fminf(fmaxf(y1, 0.f), ih); y2 = fminf(fmaxf(y2, 0.f), ih);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = lane + 32 * i;
    const float pr = e[i] / se;
    if (c < NCLS && finite && pr > thr) {
      const int slot = atomicAdd(&cand_cnt[img], 1);
      if (slot < Ccap) {
        const size_t o = (size_t)img * Ccap + slot;
        reinterpret_cast<float4*>(cand_box)[o] = make_float4(x1, y1, x2, y2);
        cand_score[o] = pr;
        cand_cls[o] = c;
        cand_canon[o] = k * NCLS + c;
      } else {
        atomicAdd(overflow, 1);
      }
    }
  }
}

__global__ void clamp_counts_kernel(int* cnt, int n, int cap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && cnt[i] > cap) cnt[i] = cap;
}

// detections from kept candidates: box / score / class from the candidate list, pred_boxes_std from the ROI row
__global__ void fastrcnn_gather_kernel(int N, int Ccap, int K, int Rcap, const int* __restrict__ keep_idx,
                                       const int* __restrict__ keep_cnt, const float* __restrict__ cand_box,
                                       const float* __restrict__ cand_score, const int* __restrict__ cand_canon,
                                       const bf16* __restrict__ pred, float* __restrict__ out_box, float* __restrict__ out_score,
                                       long long* __restrict__ out_cls, float* __restrict__ out_std, int* __restrict__ out_roi,
                                       int* __restrict__ out_cnt) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * K) return;
  const int img = t / K, k = t - img * K;
  if (k == 0) out_cnt[img] = keep_cnt[img];
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f), s = b;
  float sc = 0.f;
  long long c = 0;
  int roi = -1;
  if (k < keep_cnt[img]) {
    const int src = keep_idx[(size_t)img * K + k];
    const size_t o = (size_t)img * Ccap + src;
    b = reinterpret_cast<const float4*>(cand_box)[o];
    sc = cand_score[o];
    const int canon = cand_canon[o];
    roi = canon / NCLS;
    c = canon - roi * NCLS;
    const bf16* p = pred + ((size_t)img * Rcap + roi) * PLD + 85;
    s = make_float4(__bfloat162float(p[0]), __bfloat162float(p[1]), __bfloat162float(p[2]), __bfloat162float(p[3]));
  }
  reinterpret_cast<float4*>(out_box)[t] = b;
  out_score[t] = sc;
  out_cls[t] = c;
  reinterpret_cast<float4*>(out_std)[t] = s;
  out_roi[t] = roi;
}

// ------------------------------------------------------------------------------------ 5. small helpers
__global__ void add_f32_bf16_kernel(const float* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ o, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(a)[i];
    float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
    if (b) {
      const uint2 u = reinterpret_cast<const uint2*>(b)[i];
      y0 = __uint_as_float(u.x << 16); y1 = __uint_as_float(u.x & 0xFFFF0000u);
      y2 = __uint_as_float(u.y << 16); y3 = __uint_as_float(u.y & 0xFFFF0000u);
    }
    __nv_bfloat162 h0 = __floats2bfloat162_rn(x.x + y0, x.y + y1), h1 = __floats2bfloat162_rn(x.z + y2, x.w + y3);
    reinterpret_cast<uint2*>(o)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
  }
}

// out[n, p, q, :] = in[n, 2p, 2q, :]  ([D2] LastLevelMaxPool: max_pool2d(kernel 1, stride 2))
__global__ void subsample2x_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int N, int H, int W, int P, int Q, int C8) {
  const size_t total = (size_t)N * P * Q * C8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    size_t r = i / C8;
    const int q = (int)(r % Q); r /= Q;
    const int p = (int)(r % P);
    const int n = (int)(r / P);
    reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(in)[(((size_t)n * H + 2 * p) * W + 2 * q) * C8 + c];
  }
}
}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

// label_and_sample_proposals[_pseudo] (roi_heads.py:138-270). prop_boxes [N,Pcap,4] + prop_cnt [N]; ground truth
// [N,G,...] + gt_cnt [N]; gt_scores / gt_std NULL for the supervised branch. keys: optional uint32 [N,key_ld] indexed by
// position in [proposals | gt]; without keys the draw is hashed from seed (+ *seed_dev * 0x9E3779B1 when seed_dev, a
// device word, is given: a captured CUDA graph then draws fresh samples on every replay). Outputs are [N,Rcap,...] (fg first, ordered by key, then bg), roi_cnt [N].
extern "C" int ut2_roi_sample(int N, int Pcap, int G, int Rcap, const float* prop_boxes, const int* prop_cnt,
                              const float* gt_boxes, const long long* gt_classes, const int* gt_cnt, const float* gt_scores,
                              const float* gt_std, const unsigned int* keys, int key_ld, unsigned int seed,
                              const unsigned int* seed_dev, float pos_fraction, float iou_thr, int num_classes, int append_gt, float* roi_box,
                              long long* roi_cls, float* roi_gtbox, float* roi_conf, float* roi_std, int* roi_src,
                              int* roi_cnt, void* stream) {
  if (N <= 0) return 0;
  if (G > GMAX) return ut2_fail(-3, "roi_sample: more than 128 ground-truth slots per image");
  if (Pcap + G > SAMPLE_CAP) return ut2_fail(-4, "roi_sample: proposals + ground truth exceed 2048 per image");
  roi_sample_kernel<<<N, 1024, 0, STREAM>>>(Pcap, G, Rcap, prop_boxes, prop_cnt, gt_boxes, gt_classes, gt_cnt, gt_scores, gt_std,
                                            keys, key_ld, seed, seed_dev, (int)(Rcap * pos_fraction), iou_thr, num_classes, append_gt,
                                            roi_box, roi_cls, roi_gtbox, roi_conf, roi_std, roi_src, roi_cnt);
  return ut2_check_launch("roi_sample");
}

static int fill_roi_levels(RoiLevels& lv, int num_levels, const void* const* feats, float* const* dfeats, const int* hw,
                           const float* scales) {
  if (num_levels < 1 || num_levels > 4) return -1;
  lv.num = num_levels;
  for (int i = 0; i < 4; ++i) {
    const int j = i < num_levels ? i : num_levels - 1;
    lv.feat[i] = feats ? static_cast<const bf16*>(feats[j]) : nullptr;
    lv.dfeat[i] = dfeats ? dfeats[j] : nullptr;
    lv.H[i] = hw[2 * j]; lv.W[i] = hw[2 * j + 1]; lv.scale[i] = scales[j];
  }
  return 0;
}

// [D2] ROIPooler / [tv] roi_align. feats: HOST array of num_levels device pointers (NHWC bf16, [N, H_l, W_l, C]);
// rois [N*Rcap, 4]; out [N*Rcap, 7, 7, C] bf16. C % 8 == 0.
extern "C" int ut2_roi_align_fwd(int num_levels, const void* const* feats, const int* hw, const float* scales, int N, int C,
                                 int Rcap, const float* rois, const int* roi_cnt, void* out, void* stream) {
  RoiLevels lv;
  if (fill_roi_levels(lv, num_levels, feats, nullptr, hw, scales)) return ut2_fail(-2, "roi_align: 1..4 levels");
  if (C % 8) return ut2_fail(-3, "roi_align: C must be a multiple of 8");
  if (N * Rcap <= 0) return 0;
  roi_align_kernel<false><<<N * Rcap, 224, 0, STREAM>>>(lv, Rcap, C, rois, roi_cnt, static_cast<bf16*>(out), nullptr);
  return ut2_check_launch("roi_align_fwd");
}

// dfeats: HOST array of device pointers to fp32 [N, H_l, W_l, C] accumulators (atomically added to; caller zeroes).
extern "C" int ut2_roi_align_bwd(int num_levels, float* const* dfeats, const int* hw, const float* scales, int N, int C,
                                 int Rcap, const float* rois, const int* roi_cnt, const void* dout, void* stream) {
  RoiLevels lv;
  if (fill_roi_levels(lv, num_levels, nullptr, dfeats, hw, scales)) return ut2_fail(-2, "roi_align: 1..4 levels");
  if (C % 8) return ut2_fail(-3, "roi_align: C must be a multiple of 8");
  if (N * Rcap <= 0) return 0;
  static int merged = -1;      // UT2_ROI_BWD_MERGED=0: the per-sample kernel (A/B runs)
  if (merged < 0) { const char* e = getenv("UT2_ROI_BWD_MERGED"); merged = e ? atoi(e) : 1; }
  if (merged) roi_align_bwd_kernel<<<N * Rcap, 224, 0, STREAM>>>(lv, Rcap, C, rois, roi_cnt, static_cast<const bf16*>(dout));
  else roi_align_kernel<true><<<N * Rcap, 224, 0, STREAM>>>(lv, Rcap, C, rois, roi_cnt, nullptr, static_cast<const bf16*>(dout));
  return ut2_check_launch("roi_align_bwd");
}

// FastRCNNFocaltLossBoundaryVarOutputLayers.losses (fast_rcnn.py:834-1084). pred [N*Rcap, 96] bf16; mode 0 supervised,
// 1 unsup_data_train. acc: double[2]; losses: float[2] = {loss_cls, loss_box_reg}.
extern "C" int ut2_fastrcnn_loss_fwd(int N, int Rcap, const void* pred, const float* rois, const long long* gt_cls,
                                     const float* gt_box, const float* gt_std, const int* roi_cnt, int mode, float wx, float wy,
                                     float clamp, float gamma, float nll_w, float ts_better, float t_cert, double* acc,
                                     float* losses, void* stream) {
  if (N <= 0) return 0;
  if (mode == 1 && !gt_std) return ut2_fail(-1, "fastrcnn_loss: pseudo mode needs gt_std");
  const int Rtot = N * Rcap;
  XyxyT T{wx, wy, clamp};
  cudaMemsetAsync(acc, 0, 16, STREAM);
  fastrcnn_loss_kernel<false><<<ut2_ceil_div((long long)Rtot * 32, 256), 256, 0, STREAM>>>(
      Rtot, Rcap, static_cast<const bf16*>(pred), rois, gt_cls, gt_box, gt_std, roi_cnt, N, mode, T, gamma, nll_w, ts_better,
      t_cert, nullptr, acc, nullptr);
  fastrcnn_loss_finalize<<<1, 32, 0, STREAM>>>(acc, roi_cnt, N, losses);
  return ut2_check_launch("fastrcnn_loss_fwd");
}

extern "C" int ut2_fastrcnn_loss_bwd(int N, int Rcap, const void* pred, const float* rois, const long long* gt_cls,
                                     const float* gt_box, const float* gt_std, const int* roi_cnt, int mode, float wx, float wy,
                                     float clamp, float gamma, float nll_w, float ts_better, float t_cert, const float* gout,
                                     void* dpred, void* stream) {
  if (N <= 0) return 0;
  const int Rtot = N * Rcap;
  XyxyT T{wx, wy, clamp};
  fastrcnn_loss_kernel<true><<<ut2_ceil_div((long long)Rtot * 32, 256), 256, 0, STREAM>>>(
      Rtot, Rcap, static_cast<const bf16*>(pred), rois, gt_cls, gt_box, gt_std, roi_cnt, N, mode, T, gamma, nll_w, ts_better,
      t_cert, gout, nullptr, static_cast<bf16*>(dpred));
  return ut2_check_launch("fastrcnn_loss_bwd");
}

// First half of [D2] fast_rcnn_inference: candidates per image (capacity Ccap). cand_cnt is clamped to Ccap; overflow
// (int[1]) counts dropped candidates (0 in any realistic regime: softmax admits < 20 classes above 0.05 per ROI).
extern "C" int ut2_fastrcnn_candidates(int N, int Rcap, const void* pred, const float* rois, const int* roi_cnt,
                                       const float* image_hw, float wx, float wy, float clamp, float score_thr, int Ccap,
                                       float* cand_box, float* cand_score, int* cand_cls, int* cand_canon, int* cand_cnt,
                                       int* overflow, void* stream) {
  if (N <= 0) return 0;
  const int Rtot = N * Rcap;
  XyxyT T{wx, wy, clamp};
  cudaMemsetAsync(cand_cnt, 0, (size_t)N * 4, STREAM);
  cudaMemsetAsync(overflow, 0, 4, STREAM);
  fastrcnn_candidates_kernel<<<ut2_ceil_div((long long)Rtot * 32, 256), 256, 0, STREAM>>>(
      Rtot, Rcap, static_cast<const bf16*>(pred), rois, roi_cnt, image_hw, T, score_thr, Ccap, cand_box, cand_score, cand_cls,
      cand_canon, cand_cnt, overflow);
  clamp_counts_kernel<<<(N + 255) / 256, 256, 0, STREAM>>>(cand_cnt, N, Ccap);
  return ut2_check_launch("fastrcnn_candidates");
}

// Second half: gather the kept candidates (output of ut2_nms_batched) into [N, K, ...] detections.
extern "C" int ut2_fastrcnn_gather(int N, int Ccap, int K, int Rcap, const int* keep_idx, const int* keep_cnt,
                                   const float* cand_box, const float* cand_score, const int* cand_canon, const void* pred,
                                   float* out_box, float* out_score, long long* out_cls, float* out_std, int* out_roi,
                                   int* out_cnt, void* stream) {
  if (N * K <= 0) return 0;
  fastrcnn_gather_kernel<<<ut2_ceil_div((long long)N * K, 256), 256, 0, STREAM>>>(
      N, Ccap, K, Rcap, keep_idx, keep_cnt, cand_box, cand_score, cand_canon, static_cast<const bf16*>(pred), out_box, out_score,
      out_cls, out_std, out_roi, out_cnt);
  return ut2_check_launch("fastrcnn_gather");
}

extern "C" int ut2_add_f32_bf16(const float* a, const void* b, void* out, long long n, void* stream) {
  if (n % 4) return ut2_fail(-2, "add_f32_bf16: n must be a multiple of 4");
  if (n <= 0) return 0;
  long long g = (n / 4 + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  add_f32_bf16_kernel<<<(int)g, 256, 0, STREAM>>>(a, static_cast<const bf16*>(b), static_cast<bf16*>(out), (size_t)(n / 4));
  return ut2_check_launch("add_f32_bf16");
}

extern "C" int ut2_subsample2x_nhwc(const void* x, void* y, int N, int H, int W, int C, void* stream) {
  if (C % 8) return ut2_fail(-2, "subsample2x: C must be a multiple of 8");
  const int P = (H - 1) / 2 + 1, Q = (W - 1) / 2 + 1;
  const long long total = (long long)N * P * Q * (C / 8);
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  subsample2x_kernel<<<(int)g, 256, 0, STREAM>>>(static_cast<const bf16*>(x), static_cast<bf16*>(y), N, H, W, P, Q, C / 8);
  return ut2_check_launch("subsample2x");
}

// ------------------------------------------------------------------------------------ Box2BoxXYXYTransform (stand-alone)
// box_regression.py:36-75 / :77-129 as two elementwise kernels for the module-level API (ubteacher/modeling/box_regression.py);
// the training / inference kernels above apply the same arithmetic in registers. Round-to-nearest intrinsics in the
// reference's evaluation order (no FMA contraction): bit-identical to the torch CPU result.
__global__ void box2box_xyxy_get_deltas_kernel(const float4* __restrict__ src, const float4* __restrict__ tgt, int n, float wx, float wy,
                                               float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 s = src[i], t = tgt[i];                       // (l, d, r, u) = (x1, y1, x2, y2)
  const float w = __fadd_rn(__fsub_rn(s.z, s.x), 1.f), h = __fadd_rn(__fsub_rn(s.w, s.y), 1.f);
  float4 o;
  o.x = __fdiv_rn(__fmul_rn(wx, __fsub_rn(t.x, s.x)), w);    // dl
  o.y = __fdiv_rn(__fmul_rn(wx, __fsub_rn(t.z, s.z)), w);    // dr
  o.z = __fdiv_rn(__fmul_rn(wy, __fsub_rn(t.y, s.y)), h);    // dd
  o.w = __fdiv_rn(__fmul_rn(wy, __fsub_rn(t.w, s.w)), h);    // du
  out[i] = o;
}

__global__ void box2box_xyxy_apply_deltas_kernel(const float* __restrict__ deltas, const float4* __restrict__ boxes, int n, int k,
                                                 float wx, float wy, float clamp, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * k) return;
  const float4 b = boxes[i / k];
  const float4 d = reinterpret_cast<const float4*>(deltas)[i];
  const float w = __fsub_rn(b.z, b.x), h = __fsub_rn(b.w, b.y);          // NO +1 here (the reference's asymmetry)
  const float dl = fminf(fmaxf(__fdiv_rn(d.x, wx), -clamp), clamp), dr = fminf(fmaxf(__fdiv_rn(d.y, wx), -clamp), clamp);
  const float dd = fminf(fmaxf(__fdiv_rn(d.z, wy), -clamp), clamp), du = fminf(fmaxf(__fdiv_rn(d.w, wy), -clamp), clamp);
  float4 o;
  o.x = __fadd_rn(__fmul_rn(dl, w), b.x);
  o.y = __fadd_rn(__fmul_rn(dd, h), b.y);
  o.z = __fadd_rn(__fmul_rn(dr, w), b.z);
  o.w = __fadd_rn(__fmul_rn(du, h), b.w);
  reinterpret_cast<float4*>(out)[i] = o;
}

extern "C" int ut2_box2box_xyxy_get_deltas(const float* src_boxes, const float* target_boxes, int n, float wx, float wy, float* deltas,
                                           void* stream) {
  if (n <= 0) return 0;
  if (!src_boxes || !target_boxes || !deltas) return ut2_fail(-1, "box2box_get_deltas: null pointer");
  box2box_xyxy_get_deltas_kernel<<<(n + 255) / 256, 256, 0, STREAM>>>(reinterpret_cast<const float4*>(src_boxes),
                                                                     reinterpret_cast<const float4*>(target_boxes), n, wx, wy,
                                                                     reinterpret_cast<float4*>(deltas));
  return ut2_check_launch("box2box_get_deltas");
}

extern "C" int ut2_box2box_xyxy_apply_deltas(const float* deltas, const float* boxes, int n, int k, float wx, float wy, float clamp,
                                             float* pred_boxes, void* stream) {
  if (n <= 0 || k <= 0) return 0;
  if (!deltas || !boxes || !pred_boxes) return ut2_fail(-1, "box2box_apply_deltas: null pointer");
  box2box_xyxy_apply_deltas_kernel<<<(n * k + 255) / 256, 256, 0, STREAM>>>(deltas, reinterpret_cast<const float4*>(boxes), n, k, wx, wy,
                                                                           clamp, pred_boxes);
  return ut2_check_launch("box2box_apply_deltas");
}
