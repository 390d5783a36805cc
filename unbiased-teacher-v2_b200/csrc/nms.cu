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
  const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float sb = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter)) > thr;
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
  if (threadIdx.x < 64) s_diag[0][threadIdx.x] = (int)threadIdx.x < n ? mimg[(size_t)threadIdx.x * MW] : 0ull;
  __syncthreads();
  for (int c = 0; c < nw; ++c) {
    const int cn = min(64, n - c * 64);
    unsigned long long next_diag = 0ull;
    if (threadIdx.x < 64 && (c + 1) * 64 + (int)threadIdx.x < n)
      next_diag = __ldg(mimg + (size_t)((c + 1) * 64 + threadIdx.x) * MW + c + 1);
    const unsigned long long* diag = s_diag[c & 1];
    if (threadIdx.x == 0) {
      unsigned long long dead = remv[c];
      int nk = s_nkeep, nr = 0;
      for (int b = 0; b < cn && nk < max_keep; ++b) {
        if (!((dead >> b) & 1ull)) {
          s_rows[nr++] = (unsigned char)b;
          keep_idx[(size_t)img * max_keep + nk] = ord[c * 64 + b];
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
    if (threadIdx.x < 64) s_diag[(c + 1) & 1][threadIdx.x] = next_diag;
    __syncthreads();
  }
  if (threadIdx.x == 0) keep_cnt[img] = s_nkeep;
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
