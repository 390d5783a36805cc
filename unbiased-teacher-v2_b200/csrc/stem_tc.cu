// ResNet stem on the tensor cores: uint8 CHW image -> (x - mean)/std -> 7x7 stride-2 conv (3 -> 64) -> FrozenBN -> ReLU,
// NHWC bf16 output. Reference: pixel normalisation + ImageList padding (ubteacher/modeling/one_stage_detector.py:88-90,
// :165-167) followed by [D2] BasicStem (reached from modeling/backbone/fpn.py:59).
//
// Implicit GEMM with M = 128 output pixels (2 output rows x 64 columns), N = 64, K = 3*7*8 = 168 padded to 192
// (k = (c*7 + r)*8 + s; the s = 7 column meets a zero weight). The A tile is never read from memory as such: each
// thread builds the row of its pixel from a normalised bf16 input patch in shared memory — one 16-byte copy per
// (channel, filter row) — straight into the 128B-swizzled K-major layout that tcgen05.mma consumes; 12 MMAs
// (M128 x N64 x K16) per tile accumulate in TMEM. The patch loads of the next tile are in flight while the current
// tile is built, multiplied and stored; one launch covers the whole image batch.
// Algorithmic work: 2 * 147 * 64 FLOP per output pixel; bytes: 3 B read per 4 output pixels, 128 B written per pixel.
#include "sm100_ptx.cuh"
#include "ut2_internal.h"

namespace ut2 {

constexpr int ST_TW = 64;                      // output pixels per tile row
constexpr int ST_TR = 2;                       // output rows per tile
constexpr int ST_PIX = ST_TW * ST_TR;          // 128 output pixels per tile (UMMA M)
constexpr int ST_KB = 3;                       // K blocks of 64
constexpr int ST_CH = 21;                      // live 8-wide K chunks: one per (channel, filter row)
constexpr int ST_A_BYTES = ST_KB * ST_PIX * 128;   // 48 KiB
constexpr int ST_B_BYTES = ST_KB * 64 * 128;       // 24 KiB
constexpr int ST_PW = 2 * ST_TW + 8;           // patch width  (needs 2*63 + 8 = 134)
constexpr int ST_PR = 2 * ST_TR + 5;           // patch rows   (2*1 + 7 = 9)
constexpr int ST_PATCH = 3 * ST_PR * ST_PW;    // 3672 bf16
constexpr int ST_PATCH_BYTES = ((ST_PATCH * 2 + 127) / 128) * 128;
constexpr int ST_SMEM = 1024 + ST_A_BYTES + ST_B_BYTES + ST_PATCH_BYTES + 64 + 512;   // + scale[64] | shift[64]
constexpr int ST_PER = (ST_PATCH + 127) / 128; // patch elements per thread (29)
constexpr int ST_MAX_IMG = 32;                 // images per launch

struct StemBatch {
  const uint8_t* img[ST_MAX_IMG];
  int h[ST_MAX_IMG], w[ST_MAX_IMG];
  int n;
};

__global__ void __launch_bounds__(128)
stem_tc_kernel(const __grid_constant__ StemBatch batch, const float* __restrict__ wgt /*[7][7][3][64] fp32*/,
               const float* __restrict__ scale, const float* __restrict__ shift, float m0, float m1, float m2,
               float is0, float is1, float is2, __nv_bfloat16* __restrict__ out, int P, int Q, int tiles_q,
               int tiles_per_img) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a __shared__ pointer: LDS / STS, not generic LD / ST
  uint8_t* sA = smem;
  uint8_t* sB = smem + ST_A_BYTES;
  __nv_bfloat16* patch = reinterpret_cast<__nv_bfloat16*>(sB + ST_B_BYTES);
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(patch) + ST_PATCH_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  float* s_ss = reinterpret_cast<float*>(bar + 8);      // scale[64] | shift[64]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  s_ss[tid] = tid < 64 ? __ldg(scale + tid) : __ldg(shift + tid - 64);
  // B operand: W[n][k] bf16, K-major, 128B-swizzled, 3 k-blocks of [64 rows][128 B]. k = (c*7 + r)*8 + s: one 16-byte
  // chunk per (channel, filter row) holding the 7 taps of that row + one zero; chunks >= 21 are zero.
  for (int i = tid; i < 64 * ST_KB * 8; i += 128) {
    const int n = i / (ST_KB * 8), ch = i % (ST_KB * 8);
    const int kb = ch >> 3, c8 = ch & 7;
    const int c = ch / 7, r = ch - 7 * c;
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int s0 = 2 * e, s1 = 2 * e + 1;
      const float a = ch < ST_CH ? wgt[((r * 7 + s0) * 3 + c) * 64 + n] : 0.f;
      const float b = (ch < ST_CH && s1 < 7) ? wgt[((r * 7 + s1) * 3 + c) * 64 + n] : 0.f;
      __nv_bfloat162 hv = __floats2bfloat162_rn(a, b);
      pk[e] = *reinterpret_cast<uint32_t*>(&hv);
    }
    *reinterpret_cast<uint4*>(sB + kb * 8192 + n * 128 + ((c8 ^ (n & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  // zero the K padding chunks of A once (chunks 21..23 of each row)
  for (int i = tid; i < ST_PIX * 3; i += 128) {
    const int r = i / 3, ch = ST_CH + i % 3;
    *reinterpret_cast<uint4*>(sA + (ch >> 3) * 16384 + r * 128 + (((ch & 7) ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
  uint32_t parity = 0;
  const int num_tiles = batch.n * tiles_per_img;

  // Software pipeline: the uint8 loads of tile t+1's input patch are issued into registers before tile t's A rows are
  // built, and converted into the (by then free) shared patch while tile t's MMAs run.
  int raw[ST_PER];
  auto load_patch = [&](int t) {
    const int im = t / tiles_per_img, rem = t - im * tiles_per_img;
    const int pt = rem / tiles_q, qt = rem - pt * tiles_q;
    const int ih0 = 2 * (pt * ST_TR) - 3, iw0 = 2 * (qt * ST_TW) - 3;
    const uint8_t* img = batch.img[im];
    const int h = batch.h[im], w = batch.w[im];
#pragma unroll
    for (int it = 0; it < ST_PER; ++it) {
      const int i = it * 128 + tid;
      const int c = i / (ST_PR * ST_PW), rr = (i / ST_PW) % ST_PR, col = i % ST_PW;
      const int ih = ih0 + rr, iw = iw0 + col;
      raw[it] = -1;      // outside the image: conv padding and ImageList padding are both zeros AFTER normalisation
      if (i < ST_PATCH && ih >= 0 && ih < h && iw >= 0 && iw < w) raw[it] = __ldg(img + (size_t)c * h * w + (size_t)ih * w + iw);
    }
  };
  auto store_patch = [&]() {
#pragma unroll
    for (int it = 0; it < ST_PER; ++it) {
      const int i = it * 128 + tid;
      if (i < ST_PATCH) {
        const int c = i / (ST_PR * ST_PW);
        const float mc = c == 0 ? m0 : (c == 1 ? m1 : m2), sc_ = c == 0 ? is0 : (c == 1 ? is1 : is2);
        patch[i] = __float2bfloat16_rn(raw[it] < 0 ? 0.f : (static_cast<float>(raw[it]) - mc) * sc_);
      }
    }
  };
  int t = blockIdx.x;
  if (t < num_tiles) {
    load_patch(t);
    store_patch();
  }
  __syncthreads();
  const int orow = tid >> 6, ql = tid & 63;          // this thread's output pixel inside the tile
  for (; t < num_tiles; t += gridDim.x) {
    const int tn = t + gridDim.x;
    if (tn < num_tiles) load_patch(tn);
    // 1. this thread's pixel row of A: chunk (c, r) = patch[c][2*orow + r][2*ql .. 2*ql + 7] (the 8th value meets a
    //    zero weight), copied as four 32-bit words: lanes read consecutive words, no bank conflicts
    {
      const uint32_t* pw = reinterpret_cast<const uint32_t*>(patch) + ql;
#pragma unroll
      for (int ch = 0; ch < ST_CH; ++ch) {
        const int c = ch / 7, r = ch - 7 * c;
        const uint32_t* src = pw + ((c * ST_PR + 2 * orow + r) * ST_PW) / 2;
        *reinterpret_cast<uint4*>(sA + (ch >> 3) * 16384 + tid * 128 + (((ch & 7) ^ (tid & 7)) << 4)) =
            make_uint4(src[0], src[1], src[2], src[3]);
      }
    }
    fence_proxy_async();
    __syncthreads();          // A complete; the patch is free
    // 2. 12 MMAs by one thread (K = 192, the last 24 columns are zero on both sides)
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < ST_KB; ++kb) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = umma_smem_desc_sw128(smem_u32(sA + kb * 16384) + k * 32, 16, 1024);
          const uint64_t bd = umma_smem_desc_sw128(smem_u32(sB + kb * 8192) + k * 32, 16, 1024);
          umma_bf16(tmem, ad, bd, idesc, (kb | k) != 0);
        }
      }
      umma_commit(bar);
    }
    if (tn < num_tiles) store_patch();     // overlaps the MMAs
    mbar_wait(bar, parity);
    parity ^= 1;
    tc_fence_after();
    // 3. epilogue: lane = pixel; 64 channels -> scale/shift/ReLU -> 128 contiguous bytes (four 256-bit stores)
    {
      const int im = t / tiles_per_img, rem = t - im * tiles_per_img;
      const int pt = rem / tiles_q, qt = rem - pt * tiles_q;
      const int p = pt * ST_TR + orow, q = qt * ST_TW + ql;
      const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
      uint32_t v[4][16];
      tmem_ld_32x16(taddr, v[0]);
      tmem_ld_32x16(taddr + 16, v[1]);
      tmem_ld_32x16(taddr + 32, v[2]);
      tmem_ld_32x16(taddr + 48, v[3]);
      tmem_ld_wait();
      if (p < P && q < Q) {
        __nv_bfloat16* op = out + (((size_t)im * P + p) * Q + q) * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int n = j * 16 + 2 * i;
            const float2 sc2 = *reinterpret_cast<const float2*>(s_ss + n), sh2 = *reinterpret_cast<const float2*>(s_ss + 64 + n);
            const float a = fmaxf(fmaf(__uint_as_float(v[j][2 * i]), sc2.x, sh2.x), 0.f);
            const float b = fmaxf(fmaf(__uint_as_float(v[j][2 * i + 1]), sc2.y, sh2.y), 0.f);
            __nv_bfloat162 hv = __floats2bfloat162_rn(a, b);
            o[i] = *reinterpret_cast<uint32_t*>(&hv);
          }
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(op + j * 16), "r"(o[0]), "r"(o[1]),
                       "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                       : "memory");
        }
      }
    }
    tc_fence_before();
    __syncthreads();      // TMEM drained, A free, next patch complete
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 64);
  }
}

}  // namespace ut2

using namespace ut2;

// imgs / hs / ws: HOST arrays (N device pointers to uint8 CHW images and their sizes); out: [N, P, Q, 64] bf16.
extern "C" int ut2_stem_conv_u8_tc_batched(const void* const* imgs, const int* hs, const int* ws, int N, const float* wgt_rsck,
                                           const float* scale, const float* shift, float m0, float m1, float m2, float s0,
                                           float s1, float s2, void* out, int P, int Q, void* stream) {
  if (!imgs || !hs || !ws || !wgt_rsck || !out) return ut2_fail(-1, "stem_tc: null pointer");
  if (N <= 0) return 0;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM);
    if (e != cudaSuccess) return ut2_fail((int)e, "stem_tc: cudaFuncSetAttribute");
    set = true;
  }
  const int tiles_q = (Q + ST_TW - 1) / ST_TW, tiles_p = (P + ST_TR - 1) / ST_TR;
  const int tpi = tiles_q * tiles_p;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  for (int i0 = 0; i0 < N; i0 += ST_MAX_IMG) {
    StemBatch b;
    b.n = N - i0 < ST_MAX_IMG ? N - i0 : ST_MAX_IMG;
    for (int i = 0; i < ST_MAX_IMG; ++i) {
      const int j = i < b.n ? i0 + i : i0;
      if (!imgs[j]) return ut2_fail(-1, "stem_tc: null image");
      b.img[i] = static_cast<const uint8_t*>(imgs[j]);
      b.h[i] = hs[j]; b.w[i] = ws[j];
    }
    const int tiles = b.n * tpi;
    const int grid = tiles < 2 * sms ? tiles : 2 * sms;
    stem_tc_kernel<<<grid, 128, ST_SMEM, static_cast<cudaStream_t>(stream)>>>(
        b, wgt_rsck, scale, shift, m0, m1, m2, 1.f / s0, 1.f / s1, 1.f / s2,
        static_cast<__nv_bfloat16*>(out) + (size_t)i0 * P * Q * 64, P, Q, tiles_q, tpi);
  }
  return ut2_check_launch("stem_tc");
}

extern "C" int ut2_stem_conv_u8_tc(const void* img_chw, int h, int w, const float* wgt_rsck, const float* scale,
                                   const float* shift, float m0, float m1, float m2, float s0, float s1, float s2,
                                   void* out, int P, int Q, void* stream) {
  const void* imgs[1] = {img_chw};
  return ut2_stem_conv_u8_tc_batched(imgs, &h, &w, 1, wgt_rsck, scale, shift, m0, m1, m2, s0, s1, s2, out, P, Q, stream);
}
