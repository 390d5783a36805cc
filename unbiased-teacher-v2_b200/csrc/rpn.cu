// Region-proposal network of the Faster R-CNN half of Unbiased Teacher v2: anchor labelling + sampling, objectness /
// localisation losses (forward and backward), per-level top-k + decode of proposals. Device-resident, no host syncs.
//
// Reference (paths under /root/reference/ubteacher; [D2] = Detectron2 v0.6):
//   modeling/proposal_generator/rpn.py:21-76   PseudoLabRPN.forward (head outputs -> losses + proposals)
//   rpn.py:78-150   label_and_sample_anchors_pseudo  ([D2] pairwise_iou, Matcher(.3/.7, low-quality), subsample_labels)
//   rpn.py:153-225  losses  (BCE-with-logits weighted by the matched teacher score, [D2] _dense_box_regression_loss)
//   [D2] find_top_rpn_proposals (called at rpn.py:72-74): per-level top-k, Box2BoxTransform.apply_deltas, clip, filter
//
// Layout: the fused RPN predictor (objectness 1x1 -> 3, anchor_deltas 1x1 -> 12) writes one level-major
// [P = N * sum(H_l W_l), 16] bf16 tensor: cols 0..2 objectness of anchors a = 0..2, cols 3 + 4a + k the deltas, col 15
// padding. Anchor index inside an image: 3 * (level_off + h * W + w) + a  (the (h, w, a) order of rpn.py:35-44).
// Anchors are never materialised: (level, h, w, a) -> box is recomputed from the cell-anchor table.
#include "ut2_internal.h"
#include <cuda_bf16.h>
#include <stdint.h>

namespace {
typedef __nv_bfloat16 bf16;
constexpr int MAXL = 8;
constexpr int NA = 3;              // anchors per location
constexpr int LD = 16;             // columns of the fused predictor output
constexpr int GMAX = 128;          // ground-truth capacity per image

struct RpnLevels {
  int num;
  int H[MAXL], W[MAXL], stride[MAXL];
  int off[MAXL + 1];               // prefix of H*W
  float cell[MAXL][NA][4];         // [D2] generate_cell_anchors: (-w/2, -h/2, w/2, h/2)
};

__device__ __forceinline__ float bf(const bf16* p) { return __bfloat162float(*p); }

// anchor `a` (index inside the image) -> level, pixel, box
__device__ __forceinline__ void anchor_of(const RpnLevels& lv, int a, int& l, int& hw, int& k, float4& box) {
  const int loc = a / NA;
  k = a - loc * NA;
  l = 0;
#pragma unroll
  for (int i = 1; i < MAXL; ++i)
    if (i < lv.num && loc >= lv.off[i]) l = i;
  hw = loc - lv.off[l];
  const int h = hw / lv.W[l], w = hw - h * lv.W[l];
  const float sx = (float)(w * lv.stride[l]), sy = (float)(h * lv.stride[l]);
  box = make_float4(sx + lv.cell[l][k][0], sy + lv.cell[l][k][1], sx + lv.cell[l][k][2], sy + lv.cell[l][k][3]);
}

// [D2] pairwise_iou element: 0 where the intersection is empty
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

// ------------------------------------------------------------------------------------ 1. match (per anchor best GT)
__global__ void __launch_bounds__(256)
rpn_match_kernel(RpnLevels lv, int A, int G, const float* __restrict__ gt_boxes, const int* __restrict__ gt_cnt,
                 float* __restrict__ aval, int* __restrict__ aidx, unsigned int* __restrict__ gt_best) {
  __shared__ float4 sg[GMAX];
  __shared__ unsigned int sbest[GMAX];
  const int img = blockIdx.y;
  const int cnt = min(gt_cnt[img], G);
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    sg[i] = reinterpret_cast<const float4*>(gt_boxes)[(size_t)img * G + i];
    sbest[i] = 0u;
  }
  __syncthreads();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  float best = -1.f;
  int bi = 0;
  float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a < A) {
    int l, hw, k;
    anchor_of(lv, a, l, hw, k, box);
  }
  for (int g = 0; g < cnt; ++g) {
    const float v = a < A ? iou_pair(sg[g], box) : 0.f;
    if (v > best) { best = v; bi = g; }
    if (__any_sync(0xffffffffu, v > 0.f)) {
      float m = v;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if ((threadIdx.x & 31) == 0) atomicMax(&sbest[g], __float_as_uint(m));
    }
  }
  if (a < A) {
    aval[(size_t)img * A + a] = cnt > 0 ? best : 0.f;
    aidx[(size_t)img * A + a] = bi;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cnt; i += blockDim.x)
    if (sbest[i]) atomicMax(&gt_best[(size_t)img * G + i], sbest[i]);
}

// ------------------------------------------------------------------------------------ 2. label ([D2] Matcher)
__global__ void __launch_bounds__(256)
rpn_label_kernel(RpnLevels lv, int A, int G, const float* __restrict__ gt_boxes, const int* __restrict__ gt_cnt,
                 const float* __restrict__ aval, const unsigned int* __restrict__ gt_best, float lo_thr, float hi_thr,
                 signed char* __restrict__ labels, int* __restrict__ counts /* [N,2] = {neg, pos} */) {
  __shared__ float4 sg[GMAX];
  __shared__ float sbest[GMAX];
  __shared__ int scount[2];
  const int img = blockIdx.y;
  const int cnt = min(gt_cnt[img], G);
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    sg[i] = reinterpret_cast<const float4*>(gt_boxes)[(size_t)img * G + i];
    sbest[i] = __uint_as_float(gt_best[(size_t)img * G + i]);
  }
  if (threadIdx.x < 2) scount[threadIdx.x] = 0;
  __syncthreads();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  int lab = -2;
  if (a < A) {
    if (cnt == 0) {
      lab = 0;                                   // Matcher on an empty matrix: labels[0] (= negative) everywhere
    } else {
      const float v = aval[(size_t)img * A + a];
      lab = v < lo_thr ? 0 : (v < hi_thr ? -1 : 1);
      int l, hw, k;
      float4 box;
      anchor_of(lv, a, l, hw, k, box);
      for (int g = 0; g < cnt; ++g)              // set_low_quality_matches_: every anchor tying a GT's best IoU
        if (iou_pair(sg[g], box) == sbest[g]) lab = 1;
    }
    labels[(size_t)img * A + a] = (signed char)lab;
  }
  const unsigned int pm = __ballot_sync(0xffffffffu, lab == 1), nm = __ballot_sync(0xffffffffu, lab == 0);
  if ((threadIdx.x & 31) == 0) {
    if (nm) atomicAdd(&scount[0], __popc(nm));
    if (pm) atomicAdd(&scount[1], __popc(pm));
  }
  __syncthreads();
  if (threadIdx.x < 2 && scount[threadIdx.x]) atomicAdd(&counts[img * 2 + threadIdx.x], scount[threadIdx.x]);
}

// ------------------------------------------------------------------------------------ 3. subsample
// [D2] subsample_labels: keep n_pos = min(#pos, batch*frac) positives and n_neg = min(#neg, batch - n_pos) negatives,
// the ones with the smallest (key, index); everything else becomes -1. One CTA per image, radix select on 64 bits.
__device__ __forceinline__ unsigned long long sel_key(const uint32_t* keys, uint32_t seed, int img, int A, int a) {
  const uint32_t k = keys ? keys[(size_t)img * A + a] : hash_key(seed, (uint32_t)img, (uint32_t)a);
  return ((unsigned long long)k << 32) | (uint32_t)a;
}

constexpr int SUB_LIST = 8192;      // shared-memory capacity for the boundary bucket of the sampling select

__global__ void __launch_bounds__(1024)
rpn_subsample_kernel(int A, int batch, int max_pos, const uint32_t* __restrict__ keys, uint32_t seed,
                     const uint32_t* __restrict__ seed_dev, const int* __restrict__ counts, signed char* __restrict__ labels) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ unsigned int s_need;
  __shared__ int s_done, s_ln;
  __shared__ unsigned int slist[SUB_LIST];
  const int img = blockIdx.x;
  if (seed_dev) seed += seed_dev[0] * 0x9E3779B1u;     // device-resident draw counter (CUDA-graph replay)
  signed char* lab = labels + (size_t)img * A;
  const int npos = counts[img * 2 + 1], nneg = counts[img * 2];
  const int n_pos = min(npos, max_pos);
  const int n_neg = min(nneg, batch - n_pos);
  for (int c = 1; c >= 0; --c) {
    const int have = c ? npos : nneg, want = c ? n_pos : n_neg;
    if (have <= want) continue;                                  // keep all of this class
    if (want == 0) {
      for (int a = threadIdx.x; a < A; a += blockDim.x)
        if (lab[a] == c) lab[a] = -1;
      __syncthreads();
      continue;
    }
    unsigned long long prefix = 0, mask = 0;
    unsigned int need = want;
    bool all_in_bucket = false;
    bool listed = false;            // the boundary bucket lives in shared memory (anchor indices)
    int ln = 0;
    for (int pass = 7; pass >= 0 && !all_in_bucket; --pass) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      if (listed) {
        for (int j = threadIdx.x; j < ln; j += blockDim.x) {
          const unsigned long long k = sel_key(keys, seed, img, A, (int)slist[j]);
          if ((k & mask) == prefix) atomicAdd(&hist[(unsigned int)(k >> (8 * pass)) & 255u], 1u);
        }
      } else {
        for (int a = threadIdx.x; a < A; a += blockDim.x) {
          if (lab[a] != c) continue;
          const unsigned long long k = sel_key(keys, seed, img, A, a);
          if ((k & mask) == prefix) atomicAdd(&hist[(unsigned int)(k >> (8 * pass)) & 255u], 1u);
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned int cum = 0;
        int b = 0;
        for (; b < 255; ++b) {
          if (cum + hist[b] >= need) break;
          cum += hist[b];
        }
        s_prefix = prefix | ((unsigned long long)b << (8 * pass));
        s_need = need - cum;
        s_done = (need - cum) == hist[b];
        s_ln = 0;
      }
      __syncthreads();
      prefix = s_prefix;
      need = s_need;
      mask |= 0xFFull << (8 * pass);
      all_in_bucket = s_done != 0;
      __syncthreads();
      if (pass == 7 && !all_in_bucket) {
        // One scan settles everything outside the boundary bucket (keys are uniform hashes: the bucket holds ~1/256 of
        // the class) and moves the bucket into shared memory; the remaining passes then run on ~10^3 instead of
        // 268 569 anchors. If the bucket does not fit, the passes keep scanning the labels.
        for (int a = threadIdx.x; a < A; a += blockDim.x) {
          if (lab[a] != c) continue;
          const unsigned long long k = sel_key(keys, seed, img, A, a) & mask;
          if (k > prefix) {
            lab[a] = -1;                                  // above the bucket: never selected
          } else if (k == prefix) {
            const int t = atomicAdd(&s_ln, 1);
            if (t < SUB_LIST) slist[t] = (unsigned int)a;
          }
        }
        __syncthreads();
        if (s_ln <= SUB_LIST) {
          listed = true;
          ln = s_ln;
        }
        __syncthreads();
      }
    }
    if (listed) {         // inside the bucket: drop the keys above the final prefix
      for (int j = threadIdx.x; j < ln; j += blockDim.x) {
        const int a = (int)slist[j];
        if ((sel_key(keys, seed, img, A, a) & mask) > prefix) lab[a] = -1;
      }
      __syncthreads();
      continue;
    }
    // selected: (k & mask) < prefix, or (k & mask) == prefix (the whole remaining bucket is taken)
    for (int a = threadIdx.x; a < A; a += blockDim.x) {
      if (lab[a] != c) continue;
      const unsigned long long k = sel_key(keys, seed, img, A, a) & mask;
      if (k > prefix) lab[a] = -1;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------ 4. losses
__device__ __forceinline__ void locate_pix(const RpnLevels& lv, int N, long long p, int& l, int& img, int& hw) {
  l = 0;
#pragma unroll
  for (int i = 1; i < MAXL; ++i)
    if (i < lv.num && p >= (long long)lv.off[i] * N) l = i;
  const long long r = p - (long long)lv.off[l] * N;
  const int HW = lv.H[l] * lv.W[l];
  img = (int)(r / HW);
  hw = (int)(r - (long long)img * HW);
}

// [D2] Box2BoxTransform.get_deltas, weights (1,1,1,1)
__device__ __forceinline__ void rpn_target(const float4 a, const float4 g, float (&t)[4]) {
  const float sw = a.z - a.x, sh = a.w - a.y;
  const float sx = a.x + 0.5f * sw, sy = a.y + 0.5f * sh;
  const float tw = g.z - g.x, th = g.w - g.y;
  const float tx = g.x + 0.5f * tw, ty = g.y + 0.5f * th;
  t[0] = (tx - sx) / sw; t[1] = (ty - sy) / sh; t[2] = logf(tw / sw); t[3] = logf(th / sh);
}

// mode bwd = 0: accumulate {sum BCE, sum L1}; bwd = 1: write d(rpn_out) rows.
template <bool BWD>
__global__ void __launch_bounds__(256)
rpn_loss_kernel(RpnLevels lv, int N, int A, int G, const bf16* __restrict__ rpn_out, const signed char* __restrict__ labels,
                const int* __restrict__ aidx, const float* __restrict__ gt_boxes, const float* __restrict__ gt_scores,
                const int* __restrict__ gt_cnt, float inv_norm, const float* __restrict__ gout, double* __restrict__ acc,
                bf16* __restrict__ drpn) {
  const long long P = (long long)lv.off[lv.num] * N;
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float s_cls = 0.f, s_loc = 0.f;
  if (p < P) {
    int l, img, hw;
    locate_pix(lv, N, p, l, img, hw);
    const uint4* rowp = reinterpret_cast<const uint4*>(rpn_out + p * LD);
    const uint4 r0 = __ldg(rowp), r1 = __ldg(rowp + 1);
    const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    float x[LD];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      x[2 * i] = __uint_as_float(rw[i] << 16);
      x[2 * i + 1] = __uint_as_float(rw[i] & 0xFFFF0000u);
    }
    float d[LD];
#pragma unroll
    for (int i = 0; i < LD; ++i) d[i] = 0.f;
    const int abase = (lv.off[l] + hw) * NA;
    const int h = hw / lv.W[l], w = hw - h * lv.W[l];
    const float sx = (float)(w * lv.stride[l]), sy = (float)(h * lv.stride[l]);
    const float g_cls = BWD ? gout[0] * inv_norm : 0.f, g_loc = BWD ? gout[1] * inv_norm : 0.f;
    const int cnt = min(gt_cnt[img], G);
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      const int lab = labels[(size_t)img * A + abase + k];
      if (lab < 0) continue;
      const int gi = aidx[(size_t)img * A + abase + k];
      float wgt = 1.f;
      if (gt_scores) wgt = cnt > 0 ? gt_scores[(size_t)img * G + gi] : 0.f;
      const float z = x[k], y = (float)lab;
      if (BWD) {
        d[k] = g_cls * wgt * (1.f / (1.f + expf(-z)) - y);
      } else {
        s_cls += wgt * (fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z))));
      }
      if (lab == 1) {
        const float4 an = make_float4(sx + lv.cell[l][k][0], sy + lv.cell[l][k][1], sx + lv.cell[l][k][2], sy + lv.cell[l][k][3]);
        const float4 gb = cnt > 0 ? reinterpret_cast<const float4*>(gt_boxes)[(size_t)img * G + gi] : make_float4(0.f, 0.f, 0.f, 0.f);
        float t[4];
        rpn_target(an, gb, t);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float e = x[3 + 4 * k + j] - t[j];
          if (BWD) d[3 + 4 * k + j] = e > 0.f ? g_loc : (e < 0.f ? -g_loc : 0.f);
          else s_loc += fabsf(e);
        }
      }
    }
    if (BWD) {
      uint32_t ow[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        __nv_bfloat162 hh = __floats2bfloat162_rn(d[2 * i], d[2 * i + 1]);
        ow[i] = *reinterpret_cast<uint32_t*>(&hh);
      }
      uint4* op = reinterpret_cast<uint4*>(drpn + p * LD);
      op[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      op[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
    }
  }
  if (!BWD) {
    __shared__ float red[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s_cls += __shfl_xor_sync(0xffffffffu, s_cls, o);
      s_loc += __shfl_xor_sync(0xffffffffu, s_loc, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s_cls; red[1][threadIdx.x >> 5] = s_loc; }
    __syncthreads();
    if (threadIdx.x < 2) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
      if (s != 0.f) atomicAdd(&acc[threadIdx.x], (double)s);
    }
  }
}

__global__ void rpn_loss_finalize(const double* __restrict__ acc, float inv_norm, float* __restrict__ losses) {
  if (threadIdx.x < 2) losses[threadIdx.x] = (float)(acc[threadIdx.x] * (double)inv_norm);
}

// ------------------------------------------------------------------------------------ 5. per-level top-k + decode
// One CTA per (level, image). Selection: the K largest objectness logits (bf16 -> exact 16-bit radix select), ties to
// the smaller anchor index. Survivors are decoded with [D2] Box2BoxTransform.apply_deltas (weights 1, dw/dh clamped
// at log(1000/16)), clipped to the image, and flagged invalid (score = -inf) when non-finite or empty.
__device__ __forceinline__ unsigned int bf2ord(unsigned short u) {
  return (u & 0x8000u) ? (unsigned int)((~u) & 0xFFFFu) : (unsigned int)(u | 0x8000u);
}

constexpr unsigned int TIE_CAP = 8192;      // shared-memory list of the anchors whose logit equals the top-k threshold

// [D2] Box2BoxTransform.apply_deltas (weights 1, dw / dh clamped) + clip to the image; invalid boxes get score -inf
__device__ __forceinline__ void emit_candidate(const RpnLevels& lv, int l, int N, int img, int HW, int hw, int k, unsigned short raw,
                                               const bf16* __restrict__ rpn_out, float ih, float iw, float scale_clamp, size_t o,
                                               float* __restrict__ cand_box, float* __restrict__ cand_score,
                                               int* __restrict__ cand_canon, int* __restrict__ cand_lvl) {
  const bf16* row = rpn_out + ((size_t)lv.off[l] * N + (size_t)img * HW + hw) * LD;
  const int h = hw / lv.W[l], w = hw - h * lv.W[l];
  const float sx = (float)(w * lv.stride[l]), sy = (float)(h * lv.stride[l]);
  const float ax1 = sx + lv.cell[l][k][0], ay1 = sy + lv.cell[l][k][1];
  const float ax2 = sx + lv.cell[l][k][2], ay2 = sy + lv.cell[l][k][3];
  const float aw = ax2 - ax1, ah = ay2 - ay1;
  const float cx = ax1 + 0.5f * aw, cy = ay1 + 0.5f * ah;
  const float dx = bf(row + 3 + 4 * k), dy = bf(row + 4 + 4 * k);
  const float dw = fminf(bf(row + 5 + 4 * k), scale_clamp), dh = fminf(bf(row + 6 + 4 * k), scale_clamp);
  const float pcx = dx * aw + cx, pcy = dy * ah + cy;
  const float pw = expf(dw) * aw, ph = expf(dh) * ah;
  float x1 = pcx - 0.5f * pw, y1 = pcy - 0.5f * ph, x2 = pcx + 0.5f * pw, y2 = pcy + 0.5f * ph;
  const float score = __uint_as_float((unsigned int)raw << 16);
  bool valid = isfinite(x1) && isfinite(y1) && isfinite(x2) && isfinite(y2) && isfinite(score);
  x1 = fminf(fmaxf(x1, 0.f), iw); x2 = fminf(fmaxf(x2, 0.f), iw);
  y1 = fminf(fmaxf(y1, 0.f), ih); y2 = fminf(fmaxf(y2, 0.f), ih);
  valid = valid && (x2 - x1 > 0.f) && (y2 - y1 > 0.f);
  reinterpret_cast<float4*>(cand_box)[o] = make_float4(x1, y1, x2, y2);
  cand_score[o] = valid ? score : -INFINITY;
  cand_canon[o] = (lv.off[l] + hw) * NA + k;
  cand_lvl[o] = l;
}

__global__ void __launch_bounds__(1024)
rpn_select_decode_kernel(RpnLevels lv, int N, int K, int Mcap, const bf16* __restrict__ rpn_out,
                         const float* __restrict__ image_hw, float scale_clamp, float* __restrict__ cand_box,
                         float* __restrict__ cand_score, int* __restrict__ cand_canon, int* __restrict__ cand_lvl) {
  const int l = blockIdx.x, img = blockIdx.y;
  const int HW = lv.H[l] * lv.W[l];
  const int n = HW * NA;
  const unsigned short* base = reinterpret_cast<const unsigned short*>(rpn_out) + ((size_t)lv.off[l] * N + (size_t)img * HW) * LD;
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_need, s_cnt, s_ties;
  __shared__ unsigned int ties[TIE_CAP];
  int slot0 = 0;                                  // first output slot of this level: sum of min(HW_i * NA, K)
  for (int i = 0; i < l; ++i) slot0 += min(lv.H[i] * lv.W[i] * NA, K);
  const int take = min(n, K);
  unsigned int T = 0, idxT = 0xFFFFFFFFu;
  if (n > K) {
    unsigned int prefix = 0, mask = 0, need = K;
    for (int pass = 1; pass >= 0; --pass) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned int k = bf2ord(base[(size_t)(i / NA) * LD + (i % NA)]);
        if ((k & mask) == prefix) atomicAdd(&hist[(k >> (8 * pass)) & 255u], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned int cum = 0;
        int b = 255;
        for (; b > 0; --b) {
          if (cum + hist[b] >= need) break;
          cum += hist[b];
        }
        s_prefix = prefix | ((unsigned int)b << (8 * pass));
        s_need = need - cum;
      }
      __syncthreads();
      prefix = s_prefix;
      need = s_need;
      mask |= 0xFFu << (8 * pass);
      __syncthreads();
    }
    T = prefix;
    // `need` of the elements with key == T are taken, smallest index first. One scan emits everything above T and moves
    // the indices of the ties into shared memory; when they fit (bf16 logits tie by the hundreds, not by the tens of
    // thousands), the three index passes and the emission of the chosen ties run on that list instead of on the whole
    // level (3 instead of 6 full scans of up to 201 600 anchors by one CTA).
    if (threadIdx.x == 0) { s_cnt = 0; s_ties = 0; }
    __syncthreads();
    const float ih0 = image_hw[img * 2], iw0 = image_hw[img * 2 + 1];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int hw = i / NA, k = i - hw * NA;
      const unsigned short raw = base[(size_t)hw * LD + k];
      const unsigned int key = bf2ord(raw);
      if (key > T) {
        const unsigned int slot = atomicAdd(&s_cnt, 1u);
        emit_candidate(lv, l, N, img, HW, hw, k, raw, rpn_out, ih0, iw0, scale_clamp, (size_t)img * Mcap + slot0 + slot, cand_box,
                       cand_score, cand_canon, cand_lvl);
      } else if (key == T) {
        const unsigned int t = atomicAdd(&s_ties, 1u);
        if (t < TIE_CAP) ties[t] = (unsigned int)i;
      }
    }
    __syncthreads();
    const bool listed = s_ties <= TIE_CAP;
    const int nt = listed ? (int)s_ties : n;
    unsigned int prefix2 = 0, mask2 = 0, need2 = need;
    for (int pass = 2; pass >= 0; --pass) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      for (int j = threadIdx.x; j < nt; j += blockDim.x) {
        unsigned int k;
        if (listed) {
          k = ties[j];
        } else {
          if (bf2ord(base[(size_t)(j / NA) * LD + (j % NA)]) != T) continue;
          k = (unsigned int)j;
        }
        if ((k & mask2) == prefix2) atomicAdd(&hist[(k >> (8 * pass)) & 255u], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned int cum = 0;
        int b = 0;
        for (; b < 255; ++b) {
          if (cum + hist[b] >= need2) break;
          cum += hist[b];
        }
        s_prefix = prefix2 | ((unsigned int)b << (8 * pass));
        s_need = need2 - cum;
      }
      __syncthreads();
      prefix2 = s_prefix;
      need2 = s_need;
      mask2 |= 0xFFu << (8 * pass);
      __syncthreads();
    }
    idxT = prefix2;
    // the chosen ties (index <= idxT)
    for (int j = threadIdx.x; j < nt; j += blockDim.x) {
      int i;
      if (listed) {
        i = (int)ties[j];
      } else {
        if (bf2ord(base[(size_t)(j / NA) * LD + (j % NA)]) != T) continue;
        i = j;
      }
      if ((unsigned int)i > idxT) continue;
      const int hw = i / NA, k = i - hw * NA;
      const unsigned int slot = atomicAdd(&s_cnt, 1u);
      if (slot >= (unsigned int)take) continue;
      emit_candidate(lv, l, N, img, HW, hw, k, base[(size_t)hw * LD + k], rpn_out, ih0, iw0, scale_clamp,
                     (size_t)img * Mcap + slot0 + slot, cand_box, cand_score, cand_canon, cand_lvl);
    }
    return;
  }
  // n <= K: every anchor of the level is a candidate
  const float ih = image_hw[img * 2], iw = image_hw[img * 2 + 1];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int hw = i / NA, k = i - hw * NA;
    emit_candidate(lv, l, N, img, HW, hw, k, base[(size_t)hw * LD + k], rpn_out, ih, iw, scale_clamp, (size_t)img * Mcap + slot0 + i,
                   cand_box, cand_score, cand_canon, cand_lvl);
  }
}

__global__ void fill_int_kernel(int* p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int fill_levels(RpnLevels& lv, int num_levels, const int* hw, const int* strides, const float* cell) {
  if (num_levels < 1 || num_levels > MAXL) return -1;
  lv.num = num_levels;
  lv.off[0] = 0;
  for (int i = 0; i < num_levels; ++i) {
    lv.H[i] = hw[2 * i]; lv.W[i] = hw[2 * i + 1]; lv.stride[i] = strides[i];
    lv.off[i + 1] = lv.off[i] + hw[2 * i] * hw[2 * i + 1];
    for (int k = 0; k < NA; ++k)
      for (int j = 0; j < 4; ++j) lv.cell[i][k][j] = cell[(i * NA + k) * 4 + j];
  }
  return 0;
}
}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

// label_and_sample_anchors[_pseudo] (rpn.py:78-150). hw / strides / cell ([levels][3][4] cell anchors) are HOST arrays.
// gt_boxes [N,G,4] f32, gt_cnt [N] i32; keys: optional uint32 [N,A] sampling keys (NULL -> hashed from seed,
// plus *seed_dev * 0x9E3779B1 when the optional device word seed_dev is given: fresh draws under CUDA-graph replay).
// Outputs: labels int8 [N,A] in {-1,0,1}, matched int32 [N,A] (argmax GT), ws: float [N,A] + uint32 [N,G] + int32 [N,2].
extern "C" long long ut2_rpn_label_workspace_bytes(int N, long long A, int G) {
  return N * A * 4 + (long long)N * G * 4 + (long long)N * 8 + 1024;
}

extern "C" int ut2_rpn_label_anchors(int num_levels, const int* hw, const int* strides, const float* cell, int N, int G,
                                     const float* gt_boxes, const int* gt_cnt, const unsigned int* keys, unsigned int seed,
                                     const unsigned int* seed_dev, int batch_per_image, float pos_fraction, float lo_thr, float hi_thr, void* workspace,
                                     long long workspace_bytes, signed char* labels, int* matched, void* stream) {
  RpnLevels lv;
  if (fill_levels(lv, num_levels, hw, strides, cell)) return ut2_fail(-2, "rpn_label: bad level count");
  if (G > GMAX) return ut2_fail(-3, "rpn_label: more than 128 ground-truth slots per image");
  const int A = lv.off[lv.num] * NA;
  if (ut2_rpn_label_workspace_bytes(N, A, G) > workspace_bytes) return ut2_fail(-5, "rpn_label: workspace too small");
  char* w = static_cast<char*>(workspace);
  float* aval = reinterpret_cast<float*>(w); w += ((size_t)N * A * 4 + 255) / 256 * 256;
  unsigned int* gt_best = reinterpret_cast<unsigned int*>(w); w += ((size_t)N * G * 4 + 255) / 256 * 256;
  int* counts = reinterpret_cast<int*>(w);
  cudaMemsetAsync(gt_best, 0, ((size_t)N * G * 4 + 255) / 256 * 256 + (size_t)N * 8, STREAM);
  dim3 grid((A + 255) / 256, N);
  rpn_match_kernel<<<grid, 256, 0, STREAM>>>(lv, A, G, gt_boxes, gt_cnt, aval, matched, gt_best);
  rpn_label_kernel<<<grid, 256, 0, STREAM>>>(lv, A, G, gt_boxes, gt_cnt, aval, gt_best, lo_thr, hi_thr, labels, counts);
  rpn_subsample_kernel<<<N, 1024, 0, STREAM>>>(A, batch_per_image, (int)(batch_per_image * pos_fraction), keys, seed, seed_dev,
                                               counts, labels);
  return ut2_check_launch("rpn_label_anchors");
}

// PseudoLabRPN.losses (rpn.py:153-225). gt_scores NULL = supervised (unweighted BCE). acc: double[2] scratch,
// losses: float[2] = {loss_rpn_cls, loss_rpn_loc}; normaliser = batch_per_image * N.
extern "C" int ut2_rpn_loss_fwd(int num_levels, const int* hw, const int* strides, const float* cell, int N, int G,
                                const void* rpn_out, const signed char* labels, const int* matched, const float* gt_boxes,
                                const float* gt_scores, const int* gt_cnt, int batch_per_image, double* acc, float* losses,
                                void* stream) {
  RpnLevels lv;
  if (fill_levels(lv, num_levels, hw, strides, cell)) return ut2_fail(-2, "rpn_loss: bad level count");
  const long long P = (long long)lv.off[lv.num] * N;
  const float inv = 1.f / (float)(batch_per_image * N);
  cudaMemsetAsync(acc, 0, 16, STREAM);
  rpn_loss_kernel<false><<<ut2_ceil_div(P, 256), 256, 0, STREAM>>>(lv, N, lv.off[lv.num] * NA, G, static_cast<const bf16*>(rpn_out),
                                                                   labels, matched, gt_boxes, gt_scores, gt_cnt, inv, nullptr, acc, nullptr);
  rpn_loss_finalize<<<1, 32, 0, STREAM>>>(acc, inv, losses);
  return ut2_check_launch("rpn_loss_fwd");
}

// gout: float[2] = d(total)/d{loss_rpn_cls, loss_rpn_loc}; drpn: [P, 16] bf16, every row written.
extern "C" int ut2_rpn_loss_bwd(int num_levels, const int* hw, const int* strides, const float* cell, int N, int G,
                                const void* rpn_out, const signed char* labels, const int* matched, const float* gt_boxes,
                                const float* gt_scores, const int* gt_cnt, int batch_per_image, const float* gout, void* drpn,
                                void* stream) {
  RpnLevels lv;
  if (fill_levels(lv, num_levels, hw, strides, cell)) return ut2_fail(-2, "rpn_loss: bad level count");
  const long long P = (long long)lv.off[lv.num] * N;
  const float inv = 1.f / (float)(batch_per_image * N);
  rpn_loss_kernel<true><<<ut2_ceil_div(P, 256), 256, 0, STREAM>>>(lv, N, lv.off[lv.num] * NA, G, static_cast<const bf16*>(rpn_out),
                                                                  labels, matched, gt_boxes, gt_scores, gt_cnt, inv, gout, nullptr,
                                                                  static_cast<bf16*>(drpn));
  return ut2_check_launch("rpn_loss_bwd");
}

// First half of [D2] find_top_rpn_proposals. image_hw: device float [N,2] (h, w). Candidate slots per image:
// Mcap >= sum_l min(3 H_l W_l, pre_topk); outputs cand_box [N,Mcap,4], cand_score [N,Mcap] (-inf = dropped),
// cand_canon [N,Mcap] (anchor index, the tie-break key), cand_lvl [N,Mcap], cand_cnt [N].
extern "C" int ut2_rpn_select_decode(int num_levels, const int* hw, const int* strides, const float* cell, int N,
                                     const void* rpn_out, const float* image_hw, int pre_topk, float scale_clamp, int Mcap,
                                     float* cand_box, float* cand_score, int* cand_canon, int* cand_lvl, int* cand_cnt,
                                     void* stream) {
  RpnLevels lv;
  if (fill_levels(lv, num_levels, hw, strides, cell)) return ut2_fail(-2, "rpn_select: bad level count");
  int M = 0;
  for (int i = 0; i < lv.num; ++i) M += lv.H[i] * lv.W[i] * NA < pre_topk ? lv.H[i] * lv.W[i] * NA : pre_topk;
  if (M > Mcap) return ut2_fail(-4, "rpn_select: candidate capacity too small");
  rpn_select_decode_kernel<<<dim3(lv.num, N), 1024, 0, STREAM>>>(lv, N, pre_topk, Mcap, static_cast<const bf16*>(rpn_out), image_hw,
                                                                 scale_clamp, cand_box, cand_score, cand_canon, cand_lvl);
  // every image has exactly M candidate slots (invalid ones carry score -inf)
  fill_int_kernel<<<(N + 255) / 256, 256, 0, STREAM>>>(cand_cnt, N, M);
  return ut2_check_launch("rpn_select_decode");
}
