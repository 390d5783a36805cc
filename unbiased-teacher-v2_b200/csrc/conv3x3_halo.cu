// 3x3 / stride 1 / pad 1 convolution for the narrow trunk stages (Cout = 64 or 128: res2 / res3 conv2 of the [D2] ResNet
// bottlenecks and their data-gradients), tcgen05 + TMEM + tiled-mode TMA.
//
// Why a second kernel next to conv_igemm.cu: the im2col formulation fetches every input pixel nine times (once per filter
// tap) through L2, and with 64 / 128 output channels a tile does too little arithmetic per fetched byte: the res3 launches
// pull ~11.5 TB/s through L2 (its throughput cap is ~6300 B/clk, 10.5 TB/s at the power-capped clocks) and run at half the
// tensor rate of the 256-channel head convolutions. Here an output tile is a 2-D patch of 8 rows x 16 columns (128 pixels) and
// each 64-channel slice of its input is fetched THREE times instead of nine: one tiled TMA box of (8 + 2) rows x 16 columns
// per horizontal tap s, whose three vertical taps r are the same shared-memory tile read at row offsets r * 16 pixels =
// r * 2 KiB — 1024-byte aligned, so every tap is a plain SWIZZLE_128B K-major UMMA operand (8-row groups = 8 consecutive
// pixels of one image row, SBO = 1024). Zero padding is the box's out-of-bounds fill (tiled-mode coordinates may be negative).
//
//   stage  = [A: 10 x 16 pixels x 64 ch = 20 KiB][B: 3 taps x block_n x 64 ch]      (44 KiB at N = 64, 68 KiB at N = 128)
//   k loop = (64-channel slice, s) super-blocks, 3 taps x 4 MMAs (K = 16) each
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue (two per TMEM lane quadrant, half of the columns each);
//   two TMEM accumulators: the epilogue of tile i overlaps the MMAs of tile i + 1. Persistent over the tiles.
//
// Epilogue: + shift (FrozenBN, scale folded into the weights), optional ReLU, optional ReLU-backward mask (the data-gradient
// launches), bf16, 32-byte stores per lane. Same arithmetic as conv_fwd_kernel: fp32 accumulation over K in a different tap
// order (s-major instead of r-major), parity-tested against fp32 F.conv2d and against the im2col kernel.
#include "sm100_ptx.cuh"
#include "tmap.cuh"
#include "ut2_internal.h"
#include <stdlib.h>

namespace ut2 {
namespace {

constexpr int H3_TH = 8, H3_TW = 16;                       // output patch: 8 rows x 16 columns = UMMA M 128
constexpr int H3_A_BYTES = (H3_TH + 2) * H3_TW * 128;      // 20 KiB: (8 + 2) x 16 pixels x 64 channels
constexpr int H3_THREADS = 320;
constexpr int H3_MAX_STAGES = 5;
constexpr int H3_SMEM_LIMIT = 232448;
constexpr int H3_BAR_BYTES = 1024;

constexpr int H3_MAXL = 5;      // pyramid levels one launch can cover (level-major buffers, like conv_igemm.cu)
struct Halo3Args {
  int N, Cin, Cout;
  int block_n;             // == Cout (64, 80 or 128)
  int num_levels, num_tiles;
  int H[H3_MAXL], W[H3_MAXL], tiles_x[H3_MAXL], per_img[H3_MAXL];      // per_img = tiles_y * tiles_x
  int tile_off[H3_MAXL + 1];                                            // first patch of level l (all images)
  int row_off[H3_MAXL];                                                 // first pixel row of level l in the level-major buffers
  int stages;
  int relu;
  const float* shift;
  const __nv_bfloat16* relu_mask;
  __nv_bfloat16* out;
};

__device__ __forceinline__ void tma_load_tile_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

struct Halo3Maps {
  CUtensorMap m[H3_MAXL];
};
struct Halo3Tile { int lv, img, h0, w0; };
__device__ __forceinline__ Halo3Tile halo_tile(const Halo3Args& a, int t) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < H3_MAXL; ++i)
    if (i < a.num_levels && t >= a.tile_off[i]) l = i;
  const int rem0 = t - a.tile_off[l];
  const int img = rem0 / a.per_img[l], rem = rem0 - img * a.per_img[l];
  const int ty = rem / a.tiles_x[l], tx = rem - ty * a.tiles_x[l];
  return Halo3Tile{l, img, ty * H3_TH, tx * H3_TW};
}

__global__ void __launch_bounds__(H3_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ Halo3Maps tmaps_x, const __grid_constant__ CUtensorMap tmap_w, const Halo3Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int STAGES = a.stages;
  const int B_TAP_BYTES = a.block_n * 128;
  const int STAGE_BYTES = H3_A_BYTES + 3 * B_TAP_BYTES;
  uint8_t* bar_base = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + H3_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + H3_MAX_STAGES;     // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_shift = reinterpret_cast<float*>(bar_base + 512);       // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_sb = 3 * (a.Cin / 64);                 // (64-channel slice, s) super-blocks per tile

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmaps_x.m[0]);
    prefetch_tmap(&tmap_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 256);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();
  griddep_launch();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------ TMA producer
      uint32_t stage = 0, phase = 0;
      const uint32_t tx_bytes = H3_A_BYTES + 3 * B_TAP_BYTES;
      for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
        const Halo3Tile T = halo_tile(a, t);
        const CUtensorMap* tmap_x = &tmaps_x.m[T.lv];
        const int img = T.img, h0 = T.h0, w0 = T.w0;
        for (int c = 0; c < a.Cin; c += 64) {
          for (int s = 0; s < 3; ++s) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * STAGE_BYTES;
            mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            tma_load_tile_4d(sa, tmap_x, &full_bar[stage], c, w0 + s - 1, h0 - 1, img);
#pragma unroll
            for (int r = 0; r < 3; ++r)
              tma_load_2d(sa + H3_A_BYTES + r * B_TAP_BYTES, &tmap_w, &full_bar[stage], (r * 3 + s) * a.Cin + c, 0);
            if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------ MMA issuer
      const uint32_t idesc = umma_idesc_bf16(128, a.block_n, 0, 0);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 128;
        for (int sb = 0; sb < num_sb; ++sb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          // one descriptor pair per stage; taps and K steps only move the 14-bit start-address field (byte offset >> 4; shared
          // memory addresses stay below 2^18, so the field cannot carry): the issuing thread is a serial resource
          const uint64_t ad0 = umma_smem_desc_sw128(sa, 16, 1024);
          const uint64_t bd0 = umma_smem_desc_sw128(sa + H3_A_BYTES, 16, 1024);
          const uint32_t btap = (uint32_t)B_TAP_BYTES >> 4;
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, ad0 + (uint64_t)(r * (H3_TW * 128 / 16) + k * 2), bd0 + (uint64_t)(r * btap + k * 2), idesc,
                        (sb | r | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // -------------------------------------------------- epilogue: 8 warps, TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int cw0 = ((a.block_n / 16 + 1) / 2) * 16;    // columns of half 0: 32 / 48 / 64 of 64 / 80 / 128
    const int c0 = half * cw0;
    const int cw = half ? a.block_n - cw0 : cw0;
    const int row = quad * 32 + lane;                   // tile row = y * 16 + x
    const int y = row >> 4, x = row & 15;
    if (threadIdx.x < 64 + 128) {                       // shift vector of the (single) n-tile, staged once
      const int n = threadIdx.x - 64;
      s_shift[n] = (a.shift && n < a.Cout) ? __ldg(a.shift + n) : 0.f;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");      // the 8 epilogue warps only
    uint32_t acc = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
      const Halo3Tile T = halo_tile(a, t);
      const int h = T.h0 + y, w = T.w0 + x;
      const int H = a.H[T.lv], W = a.W[T.lv];
      const bool ok = h < H && w < W;
      const size_t orow = ((size_t)a.row_off[T.lv] + ((size_t)T.img * H + h) * W + w) * a.Cout;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 128 + c0;
      uint32_t v[2][16];
      tmem_ld_32x16(taddr, v[0]);
      const int nj16 = cw / 16;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < nj16) {
          tmem_ld_wait();
          if (j + 1 < nj16) tmem_ld_32x16(taddr + (j + 1) * 16, v[(j + 1) & 1]);
          const uint32_t(&vj)[16] = v[j & 1];
          const int cj = c0 + j * 16;
          if (ok) {
            float f[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 hv = *reinterpret_cast<const float4*>(s_shift + cj + 4 * i);
              f[4 * i] = __uint_as_float(vj[4 * i]) + hv.x;
              f[4 * i + 1] = __uint_as_float(vj[4 * i + 1]) + hv.y;
              f[4 * i + 2] = __uint_as_float(vj[4 * i + 2]) + hv.z;
              f[4 * i + 3] = __uint_as_float(vj[4 * i + 3]) + hv.w;
            }
            if (a.relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            uint32_t ow[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __nv_bfloat162 hh = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
              ow[i] = *reinterpret_cast<uint32_t*>(&hh);
            }
            if (a.relu_mask) {
              const uint4* mp = reinterpret_cast<const uint4*>(a.relu_mask + orow + cj);
              const uint4 y0 = __ldg(mp), y1 = __ldg(mp + 1);
              const uint32_t yw[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const __nv_bfloat162 mv = *reinterpret_cast<const __nv_bfloat162*>(&yw[i]);
                ow[i] &= __hgt2_mask(mv, __float2bfloat162_rn(0.f));
              }
            }
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(a.out + orow + cj), "r"(ow[0]), "r"(ow[1]),
                         "r"(ow[2]), "r"(ow[3]), "r"(ow[4]), "r"(ow[5]), "r"(ow[6]), "r"(ow[7])
                         : "memory");
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 256);
  }
}

// NHWC bf16 activation [N, H, W, C] as a 4-D tiled map, box = {64 channels, 16 columns, 8 + 2 rows, 1 image}
int make_tmap_halo(CUtensorMap* m, const void* ptr, int N, int H, int W, int C) {
  const TmapApi& api = tmap_api();
  if (!api.ok) return -100;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, H3_TW, H3_TH + 2, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = api.tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -101;
}

}  // namespace

static long long g_halo_launches = 0;

// Returns 1 when the launch was taken (0: not eligible, the caller uses the im2col kernel; < 0: error). hw: [H, W] per
// level; several levels = one launch over a level-major pyramid (the FCOS predictors: Cout = 80).
int conv3x3_halo_try(const void* x, int num_levels, const int* hw, int N, int Cin, const void* w, int Cout, const float* shift,
                     const void* relu_mask, int relu, void* y, int sm_budget, void* stream) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("UT2_HALO3"); on = e ? atoi(e) : 1; }
  static int on80 = -1;      // UT2_HALO80=0: the 80-channel predictors stay on the im2col kernel (A/B runs)
  if (on80 < 0) { const char* e = getenv("UT2_HALO80"); on80 = e ? atoi(e) : 1; }
  if (Cout == 80 && !on80) return 0;
  if (!on || (Cout != 64 && Cout != 80 && Cout != 128) || Cin % 64 || Cin > 512 || num_levels < 1 || num_levels > H3_MAXL) return 0;
  Halo3Args a;
  a.N = N; a.Cin = Cin; a.Cout = Cout; a.block_n = Cout; a.num_levels = num_levels;
  a.tile_off[0] = 0;
  long long rows = 0;
  for (int l = 0; l < H3_MAXL; ++l) {
    const int ll = l < num_levels ? l : num_levels - 1;
    a.H[l] = hw[2 * ll]; a.W[l] = hw[2 * ll + 1];
    a.tiles_x[l] = (a.W[l] + H3_TW - 1) / H3_TW;
    a.per_img[l] = ((a.H[l] + H3_TH - 1) / H3_TH) * a.tiles_x[l];
    a.row_off[l] = l < num_levels ? (int)rows : 0;
    a.tile_off[l + 1] = a.tile_off[l] + (l < num_levels ? N * a.per_img[l] : 0);
    if (l < num_levels) rows += (long long)N * a.H[l] * a.W[l];
  }
  a.num_tiles = a.tile_off[num_levels];
  if (a.num_tiles < sm_budget) return 0;               // less than one wave of patches: the im2col kernel's narrow tiles do better
  const int stage = H3_A_BYTES + 3 * Cout * 128;
  a.stages = (H3_SMEM_LIMIT - 1024 - H3_BAR_BYTES) / stage;
  if (a.stages > H3_MAX_STAGES) a.stages = H3_MAX_STAGES;
  if (a.stages < 2) return 0;
  a.relu = relu; a.shift = shift;
  a.relu_mask = static_cast<const __nv_bfloat16*>(relu_mask);
  a.out = static_cast<__nv_bfloat16*>(y);
  Halo3Maps tx;
  CUtensorMap tw;
  int rc = 0;
  for (int l = 0; l < H3_MAXL; ++l) {
    if (l < num_levels) {
      rc = make_tmap_halo(&tx.m[l], static_cast<const __nv_bfloat16*>(x) + (size_t)a.row_off[l] * Cin, N, a.H[l], a.W[l], Cin);
      if (rc) return ut2_fail(rc, "conv3x3_halo: activation tensor map encode failed");
    } else {
      tx.m[l] = tx.m[num_levels - 1];
    }
  }
  rc = make_tmap_2d_bf16(&tw, w, Cout, (uint64_t)9 * Cin, (uint64_t)9 * Cin, 64, Cout);
  if (rc) return ut2_fail(rc, "conv3x3_halo: weight tensor map encode failed");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H3_SMEM_LIMIT);
    if (e != cudaSuccess) return ut2_fail((int)e, "conv3x3_halo: cudaFuncSetAttribute");
    attr_set = true;
  }
  const int grid = a.num_tiles < sm_budget ? a.num_tiles : sm_budget;
  const size_t smem = 1024 + (size_t)a.stages * stage + H3_BAR_BYTES;
  const double M = (double)rows;
  ut2_launch_pdl(conv3x3_halo_kernel, dim3(grid), dim3(H3_THREADS), smem, static_cast<cudaStream_t>(stream),
                 ut2_est_us(2.0 * M * Cout * 9 * Cin, 2.0 * M * (Cin + Cout * (relu_mask ? 2.0 : 1.0))), tx, tw, a);
  rc = ut2_check_launch("conv3x3_halo");
  if (!rc) ++g_halo_launches;
  return rc ? rc : 1;
}

}  // namespace ut2

extern "C" long long ut2_conv3x3_halo_launches(void) { return ut2::g_halo_launches; }
