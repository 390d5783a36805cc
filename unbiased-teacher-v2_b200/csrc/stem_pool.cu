// ResNet stem + max-pool in two launches: uint8 CHW images -> (x - mean) / std -> 7x7 stride-2 conv (3 -> 64)
// -> FrozenBN -> ReLU -> 3x3 stride-2 max-pool, NHWC bf16 output [N, H/4, W/4, 64]. Reference: pixel normalisation + ImageList
// padding (ubteacher/modeling/one_stage_detector.py:88-90, :165-167), [D2] BasicStem (conv1 + max_pool2d(3, 2, 1)).
//
// The 400 x 672 x 64 stem activation (34 MB per image) never reaches HBM: 3.2 MB of pixels in, 8.6 MB + 8.6 MB of space-to-depth
// scratch, 8.6 MB of pooled output.
//
// 1. stem_s2d_kernel: normalise and SPACE-TO-DEPTH the image: x2[n][a][b][(dy*2+dx)*3 + c] = pixel(c, 2a+dy, 2b+dx), 12 live +
//    4 zero channels, bf16, zero borders (conv padding and ImageList padding are zeros AFTER normalisation). A 7x7 stride-2
//    convolution on 3 channels is a 4x4 stride-1 convolution on these 12: tap (r', s') of pixel (p, q) reads x2[p+r'-2][q+s'-2],
//    and the four taps of one filter row are 64 CONTIGUOUS bf16 (128 bytes) starting at pixel q-2.
// 2. stem_pool_tc_kernel: implicit GEMM on tcgen05, M = 128 (2 conv rows x 64 columns), N = 64, K = 4 filter rows x 64. The A
//    operand of a tile is ONE tiled TMA box [5 rows][64 windows][128 B] of a tensor map whose "window" dimension has a 32-byte
//    stride under a 128-byte inner extent (overlapping windows): k-block r' is rows (r', r'+1) of the box, no im2col copy, no
//    per-thread gather. Warp-specialised: TMA producer, MMA issuer (16 MMAs per tile, TMEM double-buffered), eight epilogue
//    warps that apply scale / shift / ReLU and park the two conv rows in an 8-row shared-memory ring, four pooling warps that
//    emit one pooled row per tile from rows (2t-1, 2t, 2t+1), up to two tiles behind the epilogue. A CTA walks DOWN a 62-column block (31 pooled columns; one warm-up tile on top).
// Same arithmetic as csrc/stem_tc.cu + ut2_maxpool3x3s2_nhwc (bf16 pixels x bf16 filter, fp32 accumulate, FrozenBN in the fp32
// epilogue, max over bf16 values): results agree up to the fp32 accumulation order.
#include "sm100_ptx.cuh"
#include "tmap.cuh"
#include "ut2_internal.h"

namespace ut2 {

constexpr int SP_PADL = 3;                     // zero pixels left of column 0 of the space-to-depth buffer
constexpr int SP_PADT = 2;                     // zero rows above row 0 (one more below)
constexpr int SP_CB = 62;                      // conv columns a column block advances (64 computed, 31 pooled)
constexpr int SP_STAGES = 3;
constexpr int SP_ROW_BYTES = 64 * 128;         // one box row: 64 windows x 128 B
constexpr int SP_STAGE_BYTES = 5 * SP_ROW_BYTES;      // 40 KiB
constexpr int SP_B_BYTES = 4 * 64 * 128;       // 4 k-blocks x [64 filters][128 B]
constexpr int SP_RING_BYTES = 8 * SP_ROW_BYTES;       // 8 conv rows x 64 pixels x 64 channels (slot = conv row & 7)
constexpr int SP_THREADS = 448;                // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quadrant), 10-13 pooling
constexpr int SP_EPI = 256;                    // epilogue threads
constexpr int SP_POOL = 128;                   // pooling threads
constexpr int SP_ROWS_PER_ITEM = 20;           // pooled rows per work item (+ 1 warm-up tile)
constexpr int SP_SMEM = 1024 + SP_STAGES * SP_STAGE_BYTES + SP_B_BYTES + SP_RING_BYTES + 1024;
constexpr int SP_MAX_IMG = 32;

struct S2dBatch {
  const uint8_t* img[SP_MAX_IMG];
  int h[SP_MAX_IMG], w[SP_MAX_IMG];
  int n;
  const int* hw_dev;       // optional DEVICE array [n][2] of (h, w): the sizes are data, not launch parameters (a captured graph of
                           // the step then serves every batch with the same padded size)
};

// one thread per (padded) space-to-depth pixel: 12 byte loads, one 32-byte store
__global__ void __launch_bounds__(256)
stem_s2d_kernel(const __grid_constant__ S2dBatch batch, float m0, float m1, float m2, float is0, float is1, float is2,
                __nv_bfloat16* __restrict__ x2, int rows, int pitch) {
  const long long total = (long long)batch.n * rows * pitch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int bc = (int)(i % pitch);
    const long long r = i / pitch;
    const int ar = (int)(r % rows), n = (int)(r / rows);
    const int a = ar - SP_PADT, b = bc - SP_PADL;
    const uint8_t* img = batch.img[n];
    const int h = batch.hw_dev ? __ldg(batch.hw_dev + 2 * n) : batch.h[n], w = batch.hw_dev ? __ldg(batch.hw_dev + 2 * n + 1) : batch.w[n];
    uint32_t o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (a >= 0 && b >= 0 && 2 * a < h && 2 * b < w) {
      float v[12];
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int y = 2 * a + dy, x = 2 * b + dx;
            const float mc = c == 0 ? m0 : (c == 1 ? m1 : m2), sc = c == 0 ? is0 : (c == 1 ? is1 : is2);
            v[(dy * 2 + dx) * 3 + c] = (y < h && x < w) ? (static_cast<float>(__ldg(img + (size_t)c * h * w + (size_t)y * w + x)) - mc) * sc : 0.f;
          }
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        __nv_bfloat162 hv = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        o[k] = *reinterpret_cast<uint32_t*>(&hv);
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(x2 + (size_t)i * 16);
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

struct StemPoolArgs {
  int N, P, Q, PP, PQ;          // conv rows / columns, pooled rows / columns
  int ncb, nrb, items;          // column blocks, row blocks per image, work items
  const float* wgt;             // [7][7][3][64] fp32 (r, s, c, k)
  const float* scale;
  const float* shift;
  __nv_bfloat16* out;           // [N, PP, PQ, 64]
};

__global__ void __launch_bounds__(SP_THREADS, 1)
stem_pool_tc_kernel(const __grid_constant__ CUtensorMap tmap, const StemPoolArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem + SP_STAGES * SP_STAGE_BYTES;
  uint8_t* ring = sB + SP_B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + SP_RING_BYTES);
  uint64_t* empty_bar = full_bar + SP_STAGES;
  uint64_t* tfull_bar = empty_bar + SP_STAGES;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;             // [2]
  uint64_t* rfull_bar = tempty_bar + 2;             // [2] conv rows of tile it are in the ring (epilogue -> pooling)
  uint64_t* rfree_bar = rfull_bar + 2;              // [2] pooling of tile it is done (pooling -> epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfree_bar + 2);
  float* s_ss = reinterpret_cast<float*>(tmem_slot + 2);      // scale[64] | shift[64]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    prefetch_tmap(&tmap);
    for (int s = 0; s < SP_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], SP_EPI);
      mbar_init(&rfull_bar[i], SP_EPI);
      mbar_init(&rfree_bar[i], SP_POOL);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  if (tid < 128) s_ss[tid] = tid < 64 ? __ldg(a.scale + tid) : __ldg(a.shift + tid - 64);
  // B operand: W[n][k], k = r' * 64 + s' * 16 + (dy * 2 + dx) * 3 + c  <->  original tap (ky, kx) = (2r' + dy - 1, 2s' + dx - 1);
  // K-major, 128B-swizzled, one [64 filters][128 B] block per filter row r'. Taps outside the 7x7 window and channels 12..15: 0.
  for (int i = tid; i < 64 * 4 * 8; i += SP_THREADS) {
    const int n = i >> 5, rp = (i >> 3) & 3, c8 = i & 7;
    uint32_t pk[4];
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      float v2[2];
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int kk = c8 * 8 + e2 * 2 + h2;
        const int sp = kk >> 4, ch = kk & 15;
        float v = 0.f;
        if (ch < 12) {
          const int dydx = ch / 3, c = ch - 3 * dydx;
          const int ky = 2 * rp + (dydx >> 1) - 1, kx = 2 * sp + (dydx & 1) - 1;
          if (ky >= 0 && ky < 7 && kx >= 0 && kx < 7) v = __ldg(a.wgt + ((ky * 7 + kx) * 3 + c) * 64 + n);
        }
        v2[h2] = v;
      }
      __nv_bfloat162 hv = __floats2bfloat162_rn(v2[0], v2[1]);
      pk[e2] = *reinterpret_cast<uint32_t*>(&hv);
    }
    *reinterpret_cast<uint4*>(sB + rp * 8192 + n * 128 + ((c8 ^ (n & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_img = a.ncb * a.nrb;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------ TMA producer: one [5 rows][64 windows][128 B] box per tile
      uint32_t stage = 0, phase = 0;
      for (int item = blockIdx.x; item < a.items; item += gridDim.x) {
        const int n = item / per_img, rem = item - n * per_img;
        const int rb = rem / a.ncb, j = rem - rb * a.ncb;
        const int t0 = rb * SP_ROWS_PER_ITEM, t1 = min(a.PP, t0 + SP_ROWS_PER_ITEM);
        const int c0 = SP_CB * j - 1;
        for (int tt = t0 > 0 ? t0 - 1 : 0; tt < t1; ++tt) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], SP_STAGE_BYTES);
          // window index of conv column q, tap s' = 0: pixel q - 2, physical q - 2 + SP_PADL; rows 2t - 2 + SP_PADT ...
          tma_load_4d(smem + stage * SP_STAGE_BYTES, &tmap, &full_bar[stage], 0, c0 - 2 + SP_PADL, 2 * tt - 2 + SP_PADT, n);
          if (++stage == SP_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------ MMA issuer: 4 filter rows x 4 K-steps of 16
      const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int item = blockIdx.x; item < a.items; item += gridDim.x) {
        const int n = item / per_img, rem = item - n * per_img;
        const int rb = rem / a.ncb;
        const int t0 = rb * SP_ROWS_PER_ITEM, t1 = min(a.PP, t0 + SP_ROWS_PER_ITEM);
        for (int tt = t0 > 0 ? t0 - 1 : 0; tt < t1; ++tt) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * SP_STAGE_BYTES);
          const uint32_t sb = smem_u32(sB);
#pragma unroll
          for (int rp = 0; rp < 4; ++rp) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = umma_smem_desc_sw128(sa + rp * SP_ROW_BYTES + k * 32, 16, 1024);    // box rows (r', r' + 1)
              const uint64_t bd = umma_smem_desc_sw128(sb + rp * 8192 + k * 32, 16, 1024);
              umma_bf16(tmem_base + acc * 64, ad, bd, idesc, (rp | k) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);
          umma_commit(&tfull_bar[acc]);
          if (++stage == SP_STAGES) { stage = 0; phase ^= 1; }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp < 10) {
    // -------------------------------------------------- epilogue: 256 threads. Two warps per TMEM lane quadrant (hardware
    // rule: warp w reads lanes 32 (w % 4) ...), each takes 32 of the 64 channels of its 32 pixels; FrozenBN + ReLU, bf16, into
    // the ring slot of the conv row. Pooling runs on its own warps, two tiles behind at most (rfull / rfree barriers).
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                // 0: channels 0..31, 1: channels 32..63
    const int m = quad * 32 + lane;                  // accumulator row: conv row (m >> 6) of the pair, column c0 + (m & 63)
    const int ri = m >> 6, px = m & 63;
    uint32_t acc = 0, acc_phase = 0, it = 0;
    for (int item = blockIdx.x; item < a.items; item += gridDim.x) {
      const int rem = item % per_img;
      const int rb = rem / a.ncb;
      const int t0 = rb * SP_ROWS_PER_ITEM, t1 = min(a.PP, t0 + SP_ROWS_PER_ITEM);
      const int first = t0 > 0 ? t0 - 1 : 0;
      for (int tt = first; tt < t1; ++tt, ++it) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 64 + half * 32;
        uint32_t v[2][16];
        tmem_ld_32x16(taddr, v[0]);
        tmem_ld_32x16(taddr + 16, v[1]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&tempty_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        // the slots of this tile's rows were last read by the pooling of tiles it-4 / it-3 (same work item) or by ANY tile of
        // the previous item: wait for pooling(it-2), and at the top of an item also for pooling(it-1)
        if (it >= 2) mbar_wait(&rfree_bar[it & 1], ((it - 2) >> 1) & 1);
        if (tt == first && it >= 1) mbar_wait(&rfree_bar[(it - 1) & 1], ((it - 1) >> 1) & 1);
        uint8_t* rrow = ring + ((2 * tt + ri) & 7) * SP_ROW_BYTES + px * 128;
#pragma unroll
        for (int jq = 0; jq < 2; ++jq) {
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = half * 32 + jq * 16 + 2 * i;
            const float2 sc2 = *reinterpret_cast<const float2*>(s_ss + c), sh2 = *reinterpret_cast<const float2*>(s_ss + 64 + c);
            const float x0 = fmaxf(fmaf(__uint_as_float(v[jq][2 * i]), sc2.x, sh2.x), 0.f);
            const float x1 = fmaxf(fmaf(__uint_as_float(v[jq][2 * i + 1]), sc2.y, sh2.y), 0.f);
            __nv_bfloat162 hv = __floats2bfloat162_rn(x0, x1);
            o[i] = *reinterpret_cast<uint32_t*>(&hv);
          }
          const int c16 = half * 4 + 2 * jq;           // 16-byte chunk of the pixel's 128-byte row
          *reinterpret_cast<uint4*>(rrow + ((c16 ^ (px & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(rrow + (((c16 + 1) ^ (px & 7)) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
        }
        mbar_arrive(&rfull_bar[it & 1]);             // release semantics: the rows are visible to the pooling warps
      }
    }
  } else {
    // -------------------------------------------------- pooling: 128 threads, pooled row tt, columns 31 j + k:
    // max over conv rows (2tt-1, 2tt, 2tt+1) x local columns (2k, 2k+1, 2k+2); 248 (column, 16-byte chunk) items per tile
    const int pt = tid - 320;                        // 0..127
    uint32_t it = 0;
    for (int item = blockIdx.x; item < a.items; item += gridDim.x) {
      const int n = item / per_img, rem = item - n * per_img;
      const int rb = rem / a.ncb, j = rem - rb * a.ncb;
      const int t0 = rb * SP_ROWS_PER_ITEM, t1 = min(a.PP, t0 + SP_ROWS_PER_ITEM);
      for (int tt = t0 > 0 ? t0 - 1 : 0; tt < t1; ++tt, ++it) {
        mbar_wait(&rfull_bar[it & 1], (it >> 1) & 1);
        if (tt >= t0) {
#pragma unroll
          for (int rep = 0; rep < 2; ++rep) {
            const int id = pt + rep * 128;
            const int k = id >> 3, ch = id & 7;
            const int pq = 31 * j + k;
            if (k < 31 && pq < a.PQ) {
              __nv_bfloat162 mx[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) mx[i] = __float2bfloat162_rn(0.f);       // post-ReLU values: 0 is the identity
#pragma unroll
              for (int dr = -1; dr <= 1; ++dr) {
                const int cr = 2 * tt + dr;
                if (cr < 0) continue;                                              // pooling pad above the image
                const uint8_t* rr = ring + (cr & 7) * SP_ROW_BYTES;
#pragma unroll
                for (int dc = 0; dc < 3; ++dc) {
                  const int lp = 2 * k + dc;
                  if (j == 0 && lp == 0) continue;                                 // conv column -1: pooling pad
                  const uint4 q4 = *reinterpret_cast<const uint4*>(rr + lp * 128 + ((ch ^ (lp & 7)) << 4));
                  const uint32_t w4[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) mx[i] = __hmax2(mx[i], *reinterpret_cast<const __nv_bfloat162*>(&w4[i]));
                }
              }
              uint4 o4;
              o4.x = *reinterpret_cast<uint32_t*>(&mx[0]); o4.y = *reinterpret_cast<uint32_t*>(&mx[1]);
              o4.z = *reinterpret_cast<uint32_t*>(&mx[2]); o4.w = *reinterpret_cast<uint32_t*>(&mx[3]);
              *reinterpret_cast<uint4*>(a.out + ((((size_t)n * a.PP + tt) * a.PQ + pq) << 6) + ch * 8) = o4;
            }
          }
        }
        mbar_arrive(&rfree_bar[it & 1]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 128);
  }
}

// space-to-depth scratch geometry for a padded batch of Hp x Wp images
struct S2dGeom {
  int H2, W2, rows, pitch, ncb, nrb, PP, PQ;
};
static S2dGeom s2d_geom(int Hp, int Wp) {
  S2dGeom g;
  g.H2 = Hp / 2; g.W2 = Wp / 2;
  g.PP = g.H2 / 2; g.PQ = g.W2 / 2;
  g.ncb = (g.PQ + 30) / 31;
  g.nrb = (g.PP + SP_ROWS_PER_ITEM - 1) / SP_ROWS_PER_ITEM;
  g.rows = g.H2 + SP_PADT + 1;
  // the last column block's box ends at window (62 (ncb - 1) - 1) - 2 + PADL + 63, each window spans 4 pixels
  int need = SP_CB * (g.ncb - 1) - 3 + SP_PADL + 63 + 4;
  if (need < g.W2 + SP_PADL + 2) need = g.W2 + SP_PADL + 2;
  g.pitch = (need + 7) / 8 * 8;
  return g;
}

}  // namespace ut2

using namespace ut2;

extern "C" long long ut2_stem_pool_workspace_bytes(int N, int Hp, int Wp) {
  if (N <= 0 || Hp <= 0 || Wp <= 0 || Hp % 4 || Wp % 4) return -1;
  const S2dGeom g = s2d_geom(Hp, Wp);
  return (long long)N * g.rows * g.pitch * 32;
}

// imgs / hs / ws: HOST arrays (N device pointers to uint8 CHW images and their sizes, each <= Hp x Wp); hw_dev: optional DEVICE
// int[N][2] with the same sizes, read by the kernel INSTEAD of hs / ws (sizes as data: CUDA-graph replay); ws: device scratch of
// ut2_stem_pool_workspace_bytes(N, Hp, Wp); out: [N, Hp/4, Wp/4, 64] bf16.
extern "C" int ut2_stem_pool_u8_batched(const void* const* imgs, const int* hs, const int* ws_, int N, const float* wgt_rsck,
                                        const float* scale, const float* shift, float m0, float m1, float m2, float s0, float s1,
                                        float s2, void* ws, long long ws_bytes, void* out, int Hp, int Wp, const int* hw_dev,
                                        void* stream) {
  if (!imgs || !hs || !ws_ || !wgt_rsck || !scale || !shift || !ws || !out) return ut2_fail(-1, "stem_pool: null pointer");
  if (N <= 0) return 0;
  if (Hp % 4 || Wp % 4) return ut2_fail(-2, "stem_pool: padded size must be a multiple of 4");
  const S2dGeom g = s2d_geom(Hp, Wp);
  if (ws_bytes < (long long)N * g.rows * g.pitch * 32) return ut2_fail(-3, "stem_pool: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* x2 = static_cast<__nv_bfloat16*>(ws);
  for (int i0 = 0; i0 < N; i0 += SP_MAX_IMG) {
    S2dBatch b;
    b.n = N - i0 < SP_MAX_IMG ? N - i0 : SP_MAX_IMG;
    for (int i = 0; i < SP_MAX_IMG; ++i) {
      const int j = i < b.n ? i0 + i : i0;
      if (!imgs[j]) return ut2_fail(-1, "stem_pool: null image");
      if (hs[j] > Hp || ws_[j] > Wp) return ut2_fail(-2, "stem_pool: image larger than the padded size");
      b.img[i] = static_cast<const uint8_t*>(imgs[j]);
      b.h[i] = hs[j]; b.w[i] = ws_[j];
    }
    b.hw_dev = hw_dev ? hw_dev + 2 * i0 : nullptr;
    const long long total = (long long)b.n * g.rows * g.pitch;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    stem_s2d_kernel<<<blocks, 256, 0, st>>>(b, m0, m1, m2, 1.f / s0, 1.f / s1, 1.f / s2, x2 + (size_t)i0 * g.rows * g.pitch * 16, g.rows,
                                            g.pitch);
  }
  // tensor map: [window element (64, stride 1)][window (pitch - 3, stride 16 elements)][row][image]
  const TmapApi& api = tmap_api();
  if (!api.ok) return ut2_fail(-100, "stem_pool: tensor map API unavailable");
  CUtensorMap tm;
  cuuint64_t dims[4] = {64, (cuuint64_t)(g.pitch - 3), (cuuint64_t)g.rows, (cuuint64_t)N};
  cuuint64_t strides[3] = {32, (cuuint64_t)g.pitch * 32, (cuuint64_t)g.rows * g.pitch * 32};
  cuuint32_t box[4] = {64, 64, 5, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = api.tiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ws, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return ut2_fail(-101, "stem_pool: tensor map encode failed (overlapping windows)");
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(stem_pool_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM);
    if (e != cudaSuccess) return ut2_fail((int)e, "stem_pool: cudaFuncSetAttribute");
    set = true;
  }
  StemPoolArgs a;
  a.N = N; a.P = g.H2; a.Q = g.W2; a.PP = g.PP; a.PQ = g.PQ;
  a.ncb = g.ncb; a.nrb = g.nrb; a.items = N * g.ncb * g.nrb;
  a.wgt = wgt_rsck; a.scale = scale; a.shift = shift;
  a.out = static_cast<__nv_bfloat16*>(out);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  sms = ut2_sm_budget(sms);
  const int grid = a.items < sms ? a.items : sms;
  stem_pool_tc_kernel<<<grid, SP_THREADS, SP_SMEM, st>>>(tm, a);
  return ut2_check_launch("stem_pool");
}
