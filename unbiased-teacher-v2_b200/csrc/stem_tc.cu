// ResNet stem on the tensor cores: uint8 CHW image -> (x - mean)/std -> 7x7 stride-2 conv (3 -> 64) -> FrozenBN -> ReLU,
// NHWC bf16 output. Reference: pixel normalisation + ImageList padding (ubteacher/modeling/one_stage_detector.py:88-90,
// :165-167) followed by [D2] BasicStem (reached from modeling/backbone/fpn.py:59).
//
// Implicit GEMM with M = 128 consecutive output pixels of one output row, N = 64, K = 7*7*3 = 147 padded to 192
// (k = (r*7 + s)*3 + c). The A tile is never read from memory as such: each thread builds the 192-wide row of its
// pixel from a normalised bf16 input patch in shared memory, straight into the 128B-swizzled K-major layout that
// tcgen05.mma consumes; 12 MMAs (M128 x N64 x K16) per tile accumulate in TMEM.
// Algorithmic work: 2 * 147 * 64 FLOP per output pixel; bytes: 3 B read per 4 output pixels, 128 B written per pixel.
#include "sm100_ptx.cuh"
#include "ut2_internal.h"

namespace ut2 {

constexpr int ST_PIX = 128;                    // output pixels per tile
constexpr int ST_KB = 3;                       // K blocks of 64
constexpr int ST_A_BYTES = ST_KB * ST_PIX * 128;   // 48 KiB
constexpr int ST_B_BYTES = ST_KB * 64 * 128;       // 24 KiB
constexpr int ST_PW = 2 * ST_PIX + 8;          // patch width (needs 2*128 + 5 = 261)
constexpr int ST_PATCH_BYTES = 3 * 7 * ST_PW * 2;  // 11088 B
constexpr int ST_SMEM = 1024 + ST_A_BYTES + ST_B_BYTES + ((ST_PATCH_BYTES + 127) / 128) * 128 + 64;

__global__ void __launch_bounds__(128)
stem_tc_kernel(const uint8_t* __restrict__ img, int h, int w, const float* __restrict__ wgt /*[7][7][3][64] fp32*/,
               const float* __restrict__ scale, const float* __restrict__ shift, float m0, float m1, float m2,
               float is0, float is1, float is2, __nv_bfloat16* __restrict__ out, int P, int Q, int tiles_per_row) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + ST_A_BYTES;
  __nv_bfloat16* patch = reinterpret_cast<__nv_bfloat16*>(sB + ST_B_BYTES);
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(patch) + ((ST_PATCH_BYTES + 127) / 128) * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  // B operand: W[n][k] bf16, K-major, 128B-swizzled, 3 k-blocks of [64 rows][128 B]; k >= 147 is zero
  for (int i = tid; i < 64 * ST_KB * 8; i += 128) {
    const int n = i / (ST_KB * 8), ch = i % (ST_KB * 8);         // ch: 16-byte chunk (8 k values) of row n
    const int kb = ch >> 3, c8 = ch & 7;
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k0 = kb * 64 + c8 * 8 + 2 * e;
      const float a = k0 < 147 ? wgt[k0 * 64 + n] : 0.f;
      const float b = k0 + 1 < 147 ? wgt[(k0 + 1) * 64 + n] : 0.f;
      __nv_bfloat162 hv = __floats2bfloat162_rn(a, b);
      pk[e] = *reinterpret_cast<uint32_t*>(&hv);
    }
    *reinterpret_cast<uint4*>(sB + kb * 8192 + n * 128 + ((c8 ^ (n & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  // zero the K padding chunks of A once (chunks 19..23 of each row: k in [152, 192)); chunk 18 is rebuilt per tile
  for (int i = tid; i < ST_PIX * 5; i += 128) {
    const int r = i / 5, ch = 19 + i % 5;
    *reinterpret_cast<uint4*>(sA + (ch >> 3) * 16384 + r * 128 + (((ch & 7) ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
  uint32_t parity = 0;
  const int num_tiles = P * tiles_per_row;
  for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    const int p = t / tiles_per_row, q0 = (t - p * tiles_per_row) * ST_PIX;
    const int ih0 = 2 * p - 3, iw0 = 2 * q0 - 3;
    // 1. normalised input patch [3][7][ST_PW] (zero outside the image: conv padding and ImageList padding)
    {
      // 21 patch rows of ST_PW bytes: all loads of a thread are issued before any is consumed (latency overlap)
      constexpr int PER = (3 * 7 * ST_PW + 127) / 128;     // 44
      int raw[PER];
#pragma unroll
      for (int it = 0; it < PER; ++it) {
        const int i = it * 128 + tid;
        const int c = i / (7 * ST_PW), r = (i / ST_PW) % 7, col = i % ST_PW;
        const int ih = ih0 + r, iw = iw0 + col;
        raw[it] = -1;
        if (i < 3 * 7 * ST_PW && ih >= 0 && ih < h && iw >= 0 && iw < w)
          raw[it] = __ldg(img + (size_t)c * h * w + (size_t)ih * w + iw);
      }
#pragma unroll
      for (int it = 0; it < PER; ++it) {
        const int i = it * 128 + tid;
        if (i < 3 * 7 * ST_PW) {
          const int c = i / (7 * ST_PW);
          const float mc = c == 0 ? m0 : (c == 1 ? m1 : m2), sc_ = c == 0 ? is0 : (c == 1 ? is1 : is2);
          patch[i] = __float2bfloat16_rn(raw[it] < 0 ? 0.f : (static_cast<float>(raw[it]) - mc) * sc_);
        }
      }
    }
    __syncthreads();
    // 2. this thread's pixel row of A: 19 chunks of 8 k-values, k = (r*7 + s)*3 + c  ->  patch[c][r][2*tid + s]
    {
      const unsigned short* pp = reinterpret_cast<const unsigned short*>(patch) + 2 * tid;
#pragma unroll
      for (int ch = 0; ch < 19; ++ch) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint32_t lo = 0, hi = 0;
          const int k0 = ch * 8 + 2 * e, k1 = k0 + 1;
          if (k0 < 147) lo = pp[(k0 % 3) * 7 * ST_PW + (k0 / 21) * ST_PW + (k0 / 3) % 7];
          if (k1 < 147) hi = pp[(k1 % 3) * 7 * ST_PW + (k1 / 21) * ST_PW + (k1 / 3) % 7];
          pk[e] = lo | (hi << 16);
        }
        *reinterpret_cast<uint4*>(sA + (ch >> 3) * 16384 + tid * 128 + (((ch & 7) ^ (tid & 7)) << 4)) =
            make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
    fence_proxy_async();
    __syncthreads();
    // 3. 12 MMAs by one thread
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
    mbar_wait(bar, parity);
    parity ^= 1;
    tc_fence_after();
    // 4. epilogue: lane = pixel row; 64 channels -> scale/shift/ReLU -> 128 contiguous bytes
    {
      const int q = q0 + warp * 32 + lane;
      const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
      uint32_t v[4][16];
      tmem_ld_32x16(taddr, v[0]);
      tmem_ld_32x16(taddr + 16, v[1]);
      tmem_ld_32x16(taddr + 32, v[2]);
      tmem_ld_32x16(taddr + 48, v[3]);
      tmem_ld_wait();
      if (q < Q) {
        uint4* op = reinterpret_cast<uint4*>(out + ((size_t)p * Q + q) * 64);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int n = j * 16 + 2 * i;
            const float a = fmaxf(fmaf(__uint_as_float(v[j][2 * i]), __ldg(scale + n), __ldg(shift + n)), 0.f);
            const float b = fmaxf(fmaf(__uint_as_float(v[j][2 * i + 1]), __ldg(scale + n + 1), __ldg(shift + n + 1)), 0.f);
            __nv_bfloat162 hv = __floats2bfloat162_rn(a, b);
            o[i] = *reinterpret_cast<uint32_t*>(&hv);
          }
          op[2 * j] = make_uint4(o[0], o[1], o[2], o[3]);
          op[2 * j + 1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();      // TMEM drained and A / patch free before the next tile
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

extern "C" int ut2_stem_conv_u8_tc(const void* img_chw, int h, int w, const float* wgt_rsck, const float* scale,
                                   const float* shift, float m0, float m1, float m2, float s0, float s1, float s2,
                                   void* out, int P, int Q, void* stream) {
  if (!img_chw || !wgt_rsck || !out) return ut2_fail(-1, "stem_tc: null pointer");
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM);
    if (e != cudaSuccess) return ut2_fail((int)e, "stem_tc: cudaFuncSetAttribute");
    set = true;
  }
  const int tpr = (Q + ST_PIX - 1) / ST_PIX;
  const int tiles = P * tpr;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = tiles < 2 * sms ? tiles : 2 * sms;
  stem_tc_kernel<<<grid, 128, ST_SMEM, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(img_chw), h, w, wgt_rsck, scale, shift, m0, m1, m2, 1.f / s0, 1.f / s1, 1.f / s2,
      static_cast<__nv_bfloat16*>(out), P, Q, tpr);
  return ut2_check_launch("stem_tc");
}
