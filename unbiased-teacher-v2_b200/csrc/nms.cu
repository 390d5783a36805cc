// Generic batched NMS on fixed-capacity candidate lists, device-resident end to end.
//
// Replaces [D2] batched_nms -> [tv] batched_nms -> [tv] nms as called from
//   [D2] find_top_rpn_proposals   (modeling/proposal_generator/rpn.py:72-74; idxs = FPN level, thr 0.7, first 1000)
//   [D2] fast_rcnn_inference      (modeling/roi_heads/fast_rcnn.py:1112-1119; idxs = class, thr 0.5, first 100)
//
// Semantics reproduced bit-exactly (fp32, round-to-nearest, no FMA contraction):
//   * candidates are visited in (score descending, tie-break key ascending) order; box i suppresses a later box j
//     of the same class iff inter / (area_i + area_j - inter) > thr, widths clamped at 0, no +1;
//   * torchvision's strategy switch: when 4 * n_valid <= trick_limit (20000 on cuda) classes are separated with the
//     coordinate trick (box + cls * (max_coord + 1), IoU evaluated on the shifted fp32 boxes), otherwise per class
//     on the raw boxes;
//   * candidates whose score is not finite are dropped (isfinite filter of both call sites).
// Output: the first max_keep survivors as indices into the candidate list, in descending score order.
#include "ut2_internal.h"
#include <stdint.h>

namespace {

__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, float thr) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float width = fmaxf(__fsub_rn(right, left), 0.f), height = fmaxf(__fsub_rn(bottom, top), 0.f);
  const float inter = __fmul_rn(width, height);
  if (!(inter > 0.f)) return false;          // 0 / union (or 0 / 0 = NaN) is never > thr
  const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float sb = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  const float uni = __fsub_rn(__fadd_rn(sa, sb), inter);
  // The exact fp32 division only decides pairs within 1e-4 of the threshold; disjoint boxes (IoU 0) and clear cases are
  // settled by the multiplication test, which cannot disagree with RN(inter / union) > thr outside that band.
  const float d = inter - thr * uni;
  if (fabsf(d) > 1e-4f * fabsf(uni)) return d > 0.f && uni > 0.f;
  return __fdiv_rn(inter, uni) > thr;
}

// order-preserving map float -> uint32 (ascending)
__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// One CTA per image: sort valid candidates by (score desc, tie asc) with an in-shared-memory bitonic network.
template <int CAP>
__global__ void __launch_bounds__(1024)
nms_sort_kernel(int M, const float* __restrict__ boxes, const float* __restrict__ scores, const int* __restrict__ tie,
                const int* __restrict__ cls, const int* __restrict__ cnt, int trick_limit, int* __restrict__ order,
                float* __restrict__ nms_box, int* __restrict__ nms_cls, int* __restrict__ n_valid) {
  extern __shared__ unsigned long long skey[];                 // CAP keys
  unsigned short* sval = reinterpret_cast<unsigned short*>(skey + CAP);
  __shared__ float smax[32];
  __shared__ int s_n;
  const int img = blockIdx.x;
  const int n_in = min(min(cnt[img], M), CAP);
  const float4* box = reinterpret_cast<const float4*>(boxes) + (size_t)img * M;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  float mx = -3.0e38f;
  int local = 0;
  int cap = 32;                    // the bitonic network spans the next power of two above the candidate count
  while (cap < n_in) cap <<= 1;
  for (int i = threadIdx.x; i < cap; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < n_in) {
      const float s = scores[(size_t)img * M + i];
      if (isfinite(s)) {
        const uint32_t t = tie ? (uint32_t)tie[(size_t)img * M + i] : (uint32_t)i;
        k = ((unsigned long long)(~f2ord(s)) << 32) | t;
        const float4 b = box[i];
        mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
        ++local;
      }
    }
    skey[i] = k;
    sval[i] = (unsigned short)i;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    local += __shfl_xor_sync(0xffffffffu, local, o);
  }
  if ((threadIdx.x & 31) == 0) {
    smax[threadIdx.x >> 5] = mx;
    atomicAdd(&s_n, local);
  }
  __syncthreads();
  mx = smax[0];
  for (int i = 1; i < 32; ++i) mx = fmaxf(mx, smax[i]);
  const int n = s_n;
  for (int k = 2; k <= cap && n_in > 1; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < cap / 2; t += blockDim.x) {     // one compare-exchange pair per thread and pass
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), ixj = i | j;
        const bool up = (i & k) == 0;
        const unsigned long long a = skey[i], b = skey[ixj];
        if ((a > b) == up) {
          skey[i] = b; skey[ixj] = a;
          const unsigned short v = sval[i]; sval[i] = sval[ixj]; sval[ixj] = v;
        }
      }
      __syncthreads();
    }
  }
  const bool trick = 4 * n <= trick_limit;
  const float offs1 = __fadd_rn(mx, 1.0f);
  if (threadIdx.x == 0) n_valid[img] = n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int src = sval[i];
    order[(size_t)img * M + i] = src;
    const int c = cls[(size_t)img * M + src];
    const float off = trick ? __fmul_rn((float)c, offs1) : 0.f;
    const float4 b = box[src];
    reinterpret_cast<float4*>(nms_box)[(size_t)img * M + i] =
        make_float4(__fadd_rn(b.x, off), __fadd_rn(b.y, off), __fadd_rn(b.z, off), __fadd_rn(b.w, off));
    nms_cls[(size_t)img * M + i] = trick ? 0 : c;
  }
}

constexpr int NMS_WPL = 4;      // mask words per lane and pass of the scan's OR phase

// mask[img][i][cb] bit j: sorted box i suppresses sorted box cb*64+j (only j > i matters)
__global__ void __launch_bounds__(64)
nms_mask_kernel(int M, int MW, const float* __restrict__ nms_box, const int* __restrict__ nms_cls,
                const int* __restrict__ n_valid, float thr, unsigned long long* __restrict__ mask) {
  const int img = blockIdx.z, rb = blockIdx.y, cb = blockIdx.x;
  const int n = n_valid[img];
  if (rb * 64 >= n || cb * 64 >= n || cb < rb) return;
  __shared__ float4 cbox[64];
  __shared__ int ccls[64];
  const float4* bx = reinterpret_cast<const float4*>(nms_box) + (size_t)img * M;
  const int* cl = nms_cls + (size_t)img * M;
  const int cn = min(64, n - cb * 64);
  if (threadIdx.x < cn) {
    cbox[threadIdx.x] = bx[cb * 64 + threadIdx.x];
    ccls[threadIdx.x] = cl[cb * 64 + threadIdx.x];
  }
  __syncthreads();
  const int i = rb * 64 + threadIdx.x;
  if (i < n) {
    const float4 a = bx[i];
    const int ac = cl[i];
    unsigned long long bits = 0;
    const int start = (rb == cb) ? threadIdx.x + 1 : 0;
    for (int j = start; j < cn; ++j)
      if (ccls[j] == ac && iou_gt(a, cbox[j], thr)) bits |= 1ull << j;
    mask[((size_t)img * M + i) * MW + cb] = bits;
  }
}

// Greedy scan in 64-box chunks: one thread resolves the dependencies inside a chunk from the diagonal mask words,
// then the whole CTA ORs the rows of that chunk's survivors into the removal bitmap: warp w takes survivors w, w+8, ...
// and its lanes the words c+1+lane, +32, ... of each row, so every thread has up to 8 x NMS_WPL independent 8-byte loads
// in flight (the kernel is one CTA per image and purely latency-bound). Stops at max_keep survivors.
__global__ void __launch_bounds__(256)
nms_scan_kernel(int M, int MW, int max_keep, const unsigned long long* __restrict__ mask, const int* __restrict__ order,
                const int* __restrict__ n_valid, int* __restrict__ keep_idx, int* __restrict__ keep_cnt) {
  extern __shared__ unsigned long long remv[];       // MW words
  __shared__ unsigned long long s_diag[2][64];      // diagonal words of chunk c / c+1 (fetched one chunk ahead)
  __shared__ int s_ord[2][64];                      // candidate indices of the chunk (no global load in the serial loop)
  __shared__ unsigned char s_rows[64];
  __shared__ int s_nrows;
  __shared__ int s_nkeep;
  const int img = blockIdx.x;
  const int n = n_valid[img];
  const int nw = (n + 63) / 64;
  for (int i = threadIdx.x; i < MW; i += blockDim.x) remv[i] = 0;
  if (threadIdx.x == 0) s_nkeep = 0;
  __syncthreads();
  const unsigned long long* mimg = mask + (size_t)img * M * MW;
  const int* ord = order + (size_t)img * M;
  if (threadIdx.x < 64) {
    s_diag[0][threadIdx.x] = (int)threadIdx.x < n ? mimg[(size_t)threadIdx.x * MW] : 0ull;
    s_ord[0][threadIdx.x] = (int)threadIdx.x < n ? ord[threadIdx.x] : 0;
  }
  __syncthreads();
  for (int c = 0; c < nw; ++c) {
    const int cn = min(64, n - c * 64);
    unsigned long long next_diag = 0ull;
    int next_ord = 0;
    if (threadIdx.x < 64 && (c + 1) * 64 + (int)threadIdx.x < n) {
      next_diag = __ldg(mimg + (size_t)((c + 1) * 64 + threadIdx.x) * MW + c + 1);
      next_ord = __ldg(ord + (c + 1) * 64 + threadIdx.x);
    }
    const unsigned long long* diag = s_diag[c & 1];
    const int* cord = s_ord[c & 1];
    if (threadIdx.x == 0) {
      unsigned long long dead = remv[c];
      int nk = s_nkeep, nr = 0;
      for (int b = 0; b < cn && nk < max_keep; ++b) {
        if (!((dead >> b) & 1ull)) {
          s_rows[nr++] = (unsigned char)b;
          keep_idx[(size_t)img * max_keep + nk] = cord[b];
          ++nk;
          dead |= diag[b];
        }
      }
      s_nrows = nr;
      s_nkeep = nk;
    }
    __syncthreads();
    if (s_nkeep >= max_keep) break;
    {
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nrows = s_nrows;
      for (int w0 = c + 1; w0 < nw; w0 += 32 * NMS_WPL) {
        unsigned long long acc[NMS_WPL];
#pragma unroll
        for (int k = 0; k < NMS_WPL; ++k) acc[k] = 0ull;
        for (int r = warp; r < nrows; r += 8) {
          const unsigned long long* row = mimg + (size_t)(c * 64 + s_rows[r]) * MW;
#pragma unroll
          for (int k = 0; k < NMS_WPL; ++k) {
            const int w = w0 + lane + 32 * k;
            if (w < nw) acc[k] |= __ldg(row + w);
          }
        }
#pragma unroll
        for (int k = 0; k < NMS_WPL; ++k) {
          const int w = w0 + lane + 32 * k;
          if (w < nw && acc[k]) atomicOr(&remv[w], acc[k]);
        }
      }
    }
    if (threadIdx.x < 64) { s_diag[(c + 1) & 1][threadIdx.x] = next_diag; s_ord[(c + 1) & 1][threadIdx.x] = next_ord; }
    __syncthreads();
  }
  if (threadIdx.x == 0) keep_cnt[img] = s_nkeep;
}

// ------------------------------------------------------------------------------------ segmented NMS (RPN proposals)
// [D2] find_top_rpn_proposals hands batched_nms the candidates of all FPN levels with idxs = level: boxes of different
// levels never interact. The candidate list is laid out level by level (segments), so each (image, level) is sorted,
// masked and scanned on its own (<= 2048 candidates: a 2048-wide sort, a 32-word mask, five CTAs per image instead of
// one 16384-wide problem) and the survivors are merged by (score desc, tie asc). Same result as the joint problem:
// suppression only acts inside a level, and each level keeps its own first max_keep survivors — a superset of what the
// joint scan would keep from it. torchvision's coordinate trick adds level * (max coordinate of the WHOLE call + 1) to the
// boxes when the call is small; that offset (and the size test) is reproduced from an image-wide prepare pass.
constexpr int SEG_MAX = 8;
constexpr int SEG_CAP = 2048;
struct SegTable {
  int S;
  int off[SEG_MAX + 1];          // segment s = slots [off[s], off[s+1]) of every image
};

__global__ void __launch_bounds__(1024)
nms_seg_prepare_kernel(int M, int total, const float* __restrict__ boxes, const float* __restrict__ scores, int trick_limit,
                       float* __restrict__ offs1, int* __restrict__ trick) {
  __shared__ float smax[32];
  __shared__ int s_n;
  const int img = blockIdx.x;
  const float4* box = reinterpret_cast<const float4*>(boxes) + (size_t)img * M;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  float mx = -3.0e38f;
  int local = 0;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    if (isfinite(scores[(size_t)img * M + i])) {
      const float4 b = box[i];
      mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
      ++local;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    local += __shfl_xor_sync(0xffffffffu, local, o);
  }
  if ((threadIdx.x & 31) == 0) {
    smax[threadIdx.x >> 5] = mx;
    atomicAdd(&s_n, local);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mx = smax[0];
    for (int i = 1; i < 32; ++i) mx = fmaxf(mx, smax[i]);
    offs1[img] = __fadd_rn(mx, 1.0f);
    trick[img] = 4 * s_n <= trick_limit;
  }
}

// One CTA per (image, segment): sort the segment's valid candidates by (score desc, tie asc); the outputs are
// segment-major ([N * S, Mseg]) so that the generic mask / scan kernels treat every segment as an "image"; order holds
// the candidate's slot inside the IMAGE's list.
__global__ void __launch_bounds__(1024)
nms_seg_sort_kernel(SegTable st, int M, int Mseg, const float* __restrict__ boxes, const float* __restrict__ scores,
                    const int* __restrict__ tie, const float* __restrict__ offs1, const int* __restrict__ trick,
                    int* __restrict__ order, float* __restrict__ nms_box, int* __restrict__ nms_cls, int* __restrict__ n_valid) {
  __shared__ unsigned long long skey[SEG_CAP];
  __shared__ unsigned short sval[SEG_CAP];
  __shared__ int s_n;
  const int img = blockIdx.x / st.S, seg = blockIdx.x - img * st.S;
  const int base = st.off[seg], n_in = st.off[seg + 1] - base;
  const float4* box = reinterpret_cast<const float4*>(boxes) + (size_t)img * M + base;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  int local = 0, cap = 32;
  while (cap < n_in) cap <<= 1;
  for (int i = threadIdx.x; i < cap; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < n_in) {
      const float s = scores[(size_t)img * M + base + i];
      if (isfinite(s)) {
        const uint32_t t = tie ? (uint32_t)tie[(size_t)img * M + base + i] : (uint32_t)(base + i);
        k = ((unsigned long long)(~f2ord(s)) << 32) | t;
        ++local;
      }
    }
    skey[i] = k;
    sval[i] = (unsigned short)i;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_n, local);
  __syncthreads();
  const int n = s_n;
  for (int k = 2; k <= cap && n_in > 1; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < cap / 2; t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), ixj = i | j;
        const bool up = (i & k) == 0;
        const unsigned long long a = skey[i], b = skey[ixj];
        if ((a > b) == up) {
          skey[i] = b; skey[ixj] = a;
          const unsigned short v = sval[i]; sval[i] = sval[ixj]; sval[ixj] = v;
        }
      }
      __syncthreads();
    }
  }
  const float off = trick[img] ? __fmul_rn((float)seg, offs1[img]) : 0.f;
  const size_t ob = (size_t)blockIdx.x * Mseg;
  if (threadIdx.x == 0) n_valid[blockIdx.x] = n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int src = sval[i];
    order[ob + i] = base + src;
    const float4 b = box[src];
    reinterpret_cast<float4*>(nms_box)[ob + i] =
        make_float4(__fadd_rn(b.x, off), __fadd_rn(b.y, off), __fadd_rn(b.z, off), __fadd_rn(b.w, off));
    nms_cls[ob + i] = 0;
  }
}

// One CTA per image: merge the S per-segment survivor lists (each in (score desc, tie asc) order) into the first max_keep
// of their union: rank of an element = its position in its own list + the number of elements of the other lists before it.
__global__ void __launch_bounds__(256)
nms_seg_merge_kernel(int S, int M, int max_keep, const float* __restrict__ scores, const int* __restrict__ tie,
                     const int* __restrict__ seg_keep, const int* __restrict__ seg_cnt, int* __restrict__ keep_idx,
                     int* __restrict__ keep_cnt) {
  extern __shared__ unsigned long long mkey[];        // [S][max_keep]
  __shared__ int cnt[SEG_MAX];
  const int img = blockIdx.x;
  if (threadIdx.x < S) cnt[threadIdx.x] = min(seg_cnt[img * S + threadIdx.x], max_keep);
  __syncthreads();
  for (int e = threadIdx.x; e < S * max_keep; e += blockDim.x) {
    const int s = e / max_keep, j = e - s * max_keep;
    unsigned long long k = ~0ull;
    if (j < cnt[s]) {
      const int slot = seg_keep[(size_t)(img * S + s) * max_keep + j];
      const uint32_t t = tie ? (uint32_t)tie[(size_t)img * M + slot] : (uint32_t)slot;
      k = ((unsigned long long)(~f2ord(scores[(size_t)img * M + slot])) << 32) | t;
    }
    mkey[e] = k;
  }
  __syncthreads();
  int total = 0;
  for (int s = 0; s < S; ++s) total += cnt[s];
  if (threadIdx.x == 0) keep_cnt[img] = min(total, max_keep);
  for (int e = threadIdx.x; e < S * max_keep; e += blockDim.x) {
    const int s = e / max_keep, j = e - s * max_keep;
    if (j >= cnt[s]) continue;
    const unsigned long long k = mkey[e];
    int rank = j;
    for (int s2 = 0; s2 < S; ++s2) {
      if (s2 == s) continue;
      const unsigned long long* lst = mkey + (size_t)s2 * max_keep;
      int lo = 0, hi = cnt[s2];                    // number of keys < k (keys are unique: the tie key is)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (lst[mid] < k) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < max_keep) keep_idx[(size_t)img * max_keep + rank] = seg_keep[(size_t)(img * S + s) * max_keep + j];
  }
}

template <typename T>
__global__ void gather_rows_kernel(int N, int M, int K, int W, const T* __restrict__ src, const int* __restrict__ idx,
                                   const int* __restrict__ cnt, T* __restrict__ dst) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)N * K * W) return;
  const int w = (int)(t % W);
  const long long r = t / W;
  const int k = (int)(r % K), img = (int)(r / K);
  dst[t] = k < cnt[img] ? src[((size_t)img * M + idx[(size_t)img * K + k]) * W + w] : T(0);
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

extern "C" long long ut2_nms_workspace_bytes(int N, int M) {
  const long long MW = (M + 63) / 64;
  return (long long)N * M * (4 + 16 + 4) + (long long)N * 4 + (long long)N * M * MW * 8 + 4096;
}

extern "C" int ut2_nms_batched(int N, int M, const float* boxes, const float* scores, const int* tie, const int* cls,
                               const int* cnt, float thr, int trick_limit, int max_keep, void* workspace,
                               long long workspace_bytes, int* keep_idx, int* keep_cnt, void* stream) {
  if (N <= 0) return 0;
  if (M <= 0 || M > 16384) return ut2_fail(-2, "nms: candidate capacity must be in [1, 16384]");
  if (!boxes || !scores || !cls || !cnt || !keep_idx || !keep_cnt) return ut2_fail(-1, "nms: null pointer");
  if (ut2_nms_workspace_bytes(N, M) > workspace_bytes) return ut2_fail(-5, "nms: workspace too small");
  const long long MW = (M + 63) / 64;
  char* w = static_cast<char*>(workspace);
  auto take = [&](long long bytes) { char* p = w; w += (bytes + 255) / 256 * 256; return p; };
  int* n_valid = reinterpret_cast<int*>(take((long long)N * 4));
  int* order = reinterpret_cast<int*>(take((long long)N * M * 4));
  int* nms_cls = reinterpret_cast<int*>(take((long long)N * M * 4));
  float* nms_box = reinterpret_cast<float*>(take((long long)N * M * 16));
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(take((long long)N * M * MW * 8));
  static bool set = false;
  if (!set) {
    cudaFuncSetAttribute(nms_sort_kernel<2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * 10);
    cudaFuncSetAttribute(nms_sort_kernel<8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 10);
    cudaFuncSetAttribute(nms_sort_kernel<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 10);
    set = true;
  }
  if (M <= 2048)
    nms_sort_kernel<2048><<<N, 1024, 2048 * 10, STREAM>>>(M, boxes, scores, tie, cls, cnt, trick_limit, order, nms_box, nms_cls, n_valid);
  else if (M <= 8192)
    nms_sort_kernel<8192><<<N, 1024, 8192 * 10, STREAM>>>(M, boxes, scores, tie, cls, cnt, trick_limit, order, nms_box, nms_cls, n_valid);
  else
    nms_sort_kernel<16384><<<N, 1024, 16384 * 10, STREAM>>>(M, boxes, scores, tie, cls, cnt, trick_limit, order, nms_box, nms_cls, n_valid);
  const int nb = (int)MW;
  nms_mask_kernel<<<dim3(nb, nb, N), 64, 0, STREAM>>>(M, (int)MW, nms_box, nms_cls, n_valid, thr, mask);
  nms_scan_kernel<<<N, 256, (size_t)MW * 8, STREAM>>>(M, (int)MW, max_keep, mask, order, n_valid, keep_idx, keep_cnt);
  return ut2_check_launch("nms_batched");
}

// Segmented form for candidate lists that are laid out class by class (the RPN: level by level): seg_off is a HOST array of
// S + 1 slot offsets inside every image's list of M slots (segment sizes <= 2048, S <= 8; every slot of a segment is a
// candidate, non-finite scores are dropped). Same keep list as ut2_nms_batched with cls = segment index.
extern "C" long long ut2_nms_segmented_workspace_bytes(int N, int S, int Mseg, int max_keep) {
  const long long NS = (long long)N * S, MW = (Mseg + 63) / 64;
  return NS * Mseg * (4 + 16 + 4) + NS * 8 + NS * Mseg * MW * 8 + NS * max_keep * 4 + (long long)N * 8 + 8192;
}

extern "C" int ut2_nms_segmented(int N, int M, int S, const int* seg_off, const float* boxes, const float* scores, const int* tie,
                                 float thr, int trick_limit, int max_keep, void* workspace, long long workspace_bytes,
                                 int* keep_idx, int* keep_cnt, void* stream) {
  if (N <= 0) return 0;
  if (!boxes || !scores || !seg_off || !keep_idx || !keep_cnt) return ut2_fail(-1, "nms_segmented: null pointer");
  if (S < 1 || S > SEG_MAX) return ut2_fail(-2, "nms_segmented: 1..8 segments");
  SegTable st;
  st.S = S;
  int Mseg = 1;
  for (int i = 0; i <= SEG_MAX; ++i) st.off[i] = seg_off[i < S ? i : S];
  for (int i = 0; i < S; ++i) {
    const int n = st.off[i + 1] - st.off[i];
    if (n < 0 || n > SEG_CAP || st.off[i + 1] > M) return ut2_fail(-2, "nms_segmented: segment sizes must be in [0, 2048]");
    if (n > Mseg) Mseg = n;
  }
  if ((size_t)S * max_keep * 8 > 200 * 1024) return ut2_fail(-2, "nms_segmented: S * max_keep too large");
  if (ut2_nms_segmented_workspace_bytes(N, S, Mseg, max_keep) > workspace_bytes) return ut2_fail(-5, "nms_segmented: workspace too small");
  const long long NS = (long long)N * S, MW = (Mseg + 63) / 64;
  char* w = static_cast<char*>(workspace);
  auto take = [&](long long bytes) { char* p = w; w += (bytes + 255) / 256 * 256; return p; };
  int* n_valid = reinterpret_cast<int*>(take(NS * 4));
  int* seg_cnt = reinterpret_cast<int*>(take(NS * 4));
  float* offs1 = reinterpret_cast<float*>(take((long long)N * 4));
  int* trick = reinterpret_cast<int*>(take((long long)N * 4));
  int* order = reinterpret_cast<int*>(take(NS * Mseg * 4));
  int* nms_cls = reinterpret_cast<int*>(take(NS * Mseg * 4));
  float* nms_box = reinterpret_cast<float*>(take(NS * Mseg * 16));
  int* seg_keep = reinterpret_cast<int*>(take(NS * max_keep * 4));
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(take(NS * Mseg * MW * 8));
  static bool set = false;
  if (!set) {
    cudaFuncSetAttribute(nms_seg_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    set = true;
  }
  nms_seg_prepare_kernel<<<N, 1024, 0, STREAM>>>(M, st.off[S], boxes, scores, trick_limit, offs1, trick);
  nms_seg_sort_kernel<<<(int)NS, 1024, 0, STREAM>>>(st, M, Mseg, boxes, scores, tie, offs1, trick, order, nms_box, nms_cls, n_valid);
  const int nb = (int)MW;
  nms_mask_kernel<<<dim3(nb, nb, (int)NS), 64, 0, STREAM>>>(Mseg, (int)MW, nms_box, nms_cls, n_valid, thr, mask);
  nms_scan_kernel<<<(int)NS, 256, (size_t)MW * 8, STREAM>>>(Mseg, (int)MW, max_keep, mask, order, n_valid, seg_keep, seg_cnt);
  nms_seg_merge_kernel<<<N, 256, (size_t)S * max_keep * 8, STREAM>>>(S, M, max_keep, scores, tie, seg_keep, seg_cnt, keep_idx, keep_cnt);
  return ut2_check_launch("nms_segmented");
}

// dst[img, k, :] = src[img, idx[img, k], :] for k < cnt[img], zero otherwise. elem_bytes in {4, 8}.
extern "C" int ut2_gather_rows(int N, int M, int K, int W, int elem_bytes, const void* src, const int* idx, const int* cnt,
                               void* dst, void* stream) {
  const long long total = (long long)N * K * W;
  if (total <= 0) return 0;
  const int g = ut2_ceil_div(total, 256);
  if (elem_bytes == 4)
    gather_rows_kernel<float><<<g, 256, 0, STREAM>>>(N, M, K, W, static_cast<const float*>(src), idx, cnt, static_cast<float*>(dst));
  else if (elem_bytes == 8)
    gather_rows_kernel<long long><<<g, 256, 0, STREAM>>>(N, M, K, W, static_cast<const long long*>(src), idx, cnt,
                                                         static_cast<long long*>(dst));
  else
    return ut2_fail(-2, "gather_rows: elem_bytes must be 4 or 8");
  return ut2_check_launch("gather_rows");
}
