// Implicit-GEMM convolution on the sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the cuDNN conv2d calls the reference reaches through Detectron2 / torch:
//   ResNet bottlenecks + FPN          (ubteacher/modeling/backbone/fpn.py:59-78 -> [D2] ResNet/FPN)
//   FCOS towers + prediction convs    (ubteacher/modeling/fcos/fcos.py:248-304, :338-376)
//
// Layouts (all bf16, fp32 accumulate):
//   activations  NHWC  == row-major [M = N*P*Q, C]
//   weights      [Cout, R, S, Cin]  == row-major [Cout, K = R*S*Cin]   (K-major B operand)
// Forward:  Y[m, n] = sum_k im2col(X)[m, k] * W[n, k]   -> epilogue: *scale[n] + shift[n] (+res) (relu)
// Dgrad  :  the same kernel on dY with the flipped/transposed filter (host packs it).
// Wgrad  :  dW[n, tap, c] = sum_m dY[m, n] * im2col(X)[m, tap, c]  (both operands MN-major,
//           reduction over pixels, split-K with fp32 red.global.add).
//
// The A operand never exists in memory: each k-block is one filter tap x 64 input channels and is
// fetched by one im2col-mode TMA load (hardware halo zero-fill, stride handled by the descriptor).
#include "sm100_ptx.cuh"
#include "tmap.cuh"
#include "ut2_internal.h"
#include <stdlib.h>

namespace ut2 {

constexpr int BM = 128;          // output pixels per tile (UMMA M)
constexpr int BK = 64;           // K elements per pipeline stage (128 B rows, SWIZZLE_128B)
constexpr int MAX_STAGES = 8;
constexpr int A_BYTES = BM * BK * 2;        // 16 KiB
constexpr int SMEM_LIMIT = 232448;          // 227 KiB of dynamic shared memory per CTA on sm_100a
constexpr int BAR_BYTES = 3072;             // 1 KiB mbarriers + TMEM slot | 1 KiB scale[256] | 1 KiB shift[256]
constexpr int EPI_TILE_BYTES = 32 * 128;    // one 32-row x 64-col bf16 staging tile, 128B-swizzled
constexpr int NUM_THREADS = 576;            // warp0 TMA, warp1 MMA, warps2-17 epilogue (four per TMEM lane quadrant)
constexpr int TMEM_COLS = 512;              // 2 accumulator buffers x 256 fp32 columns

// shared memory: [stages x (16 KiB A + block_n x 128 B of B)][barriers, scale, shift][16 warps x 1 aux tile of 4 KiB
// (only with aux)]. The ring is as deep as the 227 KiB allow (<= 8): narrow-N tiles move few bytes per stage, and with
// only 4 stages in flight they are bound by TMA latency, not by bandwidth or the tensor pipe.
__host__ __device__ constexpr int stage_bytes(int block_n) { return A_BYTES + block_n * BK * 2; }
// aux_tiles: 0 (no residual / mask tile), 16 (one staging tile per epilogue warp) or 32 (double-buffered). pad: 1024 bytes
// of slack for the run-time 1024-byte alignment of the buffer; the double-buffered 256-wide case fits the 227 KiB only
// without it (2 x 48 KiB stages + 3 KiB + 128 KiB = 227 KiB exactly) and then REQUIRES an aligned base (checked).
constexpr int smem_bytes(int stages, int block_n, int aux_tiles, int pad) {
  return pad + stages * stage_bytes(block_n) + BAR_BYTES + aux_tiles * EPI_TILE_BYTES;
}
inline int pick_stages(int block_n, int aux_tiles, int pad) {
  int s = (SMEM_LIMIT - pad - BAR_BYTES - aux_tiles * EPI_TILE_BYTES) / stage_bytes(block_n);
  return s > MAX_STAGES ? MAX_STAGES : s;
}

constexpr int MAX_LV = 5;          // pyramid levels one launch can cover (shared-weight head convs)

// Per-level geometry of a launch. An ordinary convolution is the 1-level case. Levels are laid back to back in one
// "level-major" buffer: level l owns rows [row_off[l], row_off[l] + M[l]) of the [sum M, C] activation matrix and
// m-tiles [tile_off[l], tile_off[l+1]) — tiles never straddle two levels (the last tile of a level is ragged).
struct LevelTable {
  int num;
  int tile_off[MAX_LV + 1];
  int row_off[MAX_LV];
  int M[MAX_LV], P[MAX_LV], Q[MAX_LV];
};
struct TmapSet {
  CUtensorMap m[MAX_LV];
};

__device__ __forceinline__ int level_of(const LevelTable& lt, int tile) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < MAX_LV; ++i)
    if (i < lt.num && tile >= lt.tile_off[i]) l = i;
  return l;
}

struct ConvFwdArgs {
  int M, Cout, ldo;
  int block_n, n_tiles, m_tiles;
  LevelTable lt;
  int P, Q, stride, pad;
  int R, S, Cin;
  int relu;
  int stages;                       // operand ring depth (pick_stages)
  int aux_dbl;                      // 1: two aux staging tiles per epilogue warp (memory-bound convs: see the epilogue)
  int no_pad;                       // 1: the dynamic shared buffer must already be 1024-byte aligned
  int aux_kind;                     // 0 none, 1 residual tile via TMA, 2 relu-mask tile via TMA, 3 both (two tiles per warp)
  int manual;                       // 1: epilogue with plain loads/stores (res_up2, Cout < 64)
  int out_tma;                      // 1: output tiles staged in shared memory and written by TMA stores (memory-bound 1x1 convs);
                                    // 2: staged IN PLACE in the aux tile the lane has just consumed (no extra shared memory)
  int aux_bytes;                    // bytes of aux staging in front of the output staging tiles
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  int ldr;
  int res_up2;                      // residual is [N, P/2, Q/2, Cout]: nearest-2x upsampled on the fly (FPN top-down)
  const __nv_bfloat16* relu_mask;   // optional [M, Cout]: output zeroed where mask <= 0 (ReLU backward fused in dgrad)
  __nv_bfloat16* out;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 16-byte chunk `c` of row `r` inside a 128B-swizzled [rows][128 B] tile (same pattern TMA SWIZZLE_128B uses)
__device__ __forceinline__ uint32_t swz(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }
// zero the bf16 halves of `o` whose mask element is not > 0 (sign set or magnitude zero)
__device__ __forceinline__ uint32_t mask_bf16x2(uint32_t o, uint32_t m) {
  const __nv_bfloat162 mv = *reinterpret_cast<const __nv_bfloat162*>(&m);
  return o & __hgt2_mask(mv, __float2bfloat162_rn(0.f));      // one HSET2.BM: 0xFFFF per half whose mask element is > 0
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_fwd_kernel(const __grid_constant__ TmapSet tmaps_x, const __grid_constant__ CUtensorMap tmap_w,
                const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_aux,
                const __grid_constant__ CUtensorMap tmap_aux2, const ConvFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a __shared__ pointer: LDS / STS, not generic LD / ST
  if (a.no_pad && smem != smem_raw) __trap();      // the launch reserved no alignment slack (see smem_bytes)
  const int STAGES = a.stages;
  const int STAGE_BYTES = stage_bytes(a.block_n);
  uint8_t* bar_base = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]
  uint64_t* aux_bar = tempty_bar + 2;             // [32] two per epilogue warp (the second one only with aux_dbl)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + 32);
  float* s_scale = reinterpret_cast<float*>(bar_base + 1024);
  float* s_shift = reinterpret_cast<float*>(bar_base + 2048);
  uint8_t* aux_stage = bar_base + BAR_BYTES;                 // 16 warps x 4 KiB (aux_kind != 0 only)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = a.m_tiles * a.n_tiles;
  const int c_chunks = (a.Cin + BK - 1) / BK;   // a ragged last chunk is zero-filled by TMA (OOB channels)
  const int num_kb = a.R * a.S * c_chunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmaps_x.m[0]);
    prefetch_tmap(&tmap_w);
    if (a.aux_kind) prefetch_tmap(&tmap_aux);
    if (a.aux_kind == 3) prefetch_tmap(&tmap_aux2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 512);
    }
    for (int i = 0; i < 32; ++i) mbar_init(&aux_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();        // everything above overlapped the previous kernel's tail (programmatic dependent launch)
  griddep_launch();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------ TMA producer
      uint32_t stage = 0, phase = 0;
      const uint32_t tx_bytes = A_BYTES + a.block_n * BK * 2;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int n_tile = t % a.n_tiles, m_tile = t / a.n_tiles;
        const int lv = level_of(a.lt, m_tile);
        const CUtensorMap* tmap_x = &tmaps_x.m[lv];
        const int m0 = (m_tile - a.lt.tile_off[lv]) * BM;
        const int Q = a.lt.Q[lv], PQ = a.lt.P[lv] * Q;
        const int img = m0 / PQ, rem = m0 - img * PQ;
        const int p0 = rem / Q, q0 = rem - p0 * Q;
        const int w0 = q0 * a.stride - a.pad, h0 = p0 * a.stride - a.pad;
        // The producer is ONE thread: its per-k-block instruction count bounds the rate of narrow-N tiles (the tensor
        // pipe needs only 32 / 64 cycles per MMA at N = 64 / 128), so the (tap, channel) walk is incremental — no
        // divisions in the loop.
        int r = 0, sft = 0, c0 = 0, kcol = 0, tapcol = 0;
        const int ncol = n_tile * a.block_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          tma_load_im2col_4d(sa, tmap_x, &full_bar[stage], c0, w0, h0, img, (uint16_t)sft,
                             (uint16_t)r);
          tma_load_2d(sa + A_BYTES, &tmap_w, &full_bar[stage], kcol, ncol);
          if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1; }
          c0 += BK;
          kcol += BK;
          if (c0 >= a.Cin) {
            c0 = 0;
            tapcol += a.Cin;
            kcol = tapcol;
            if (++sft == a.S) { sft = 0; ++r; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------ MMA issuer
      const uint32_t idesc = umma_idesc_bf16(BM, a.block_n, 0, 0);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = umma_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t bd = umma_smem_desc_sw128(sb + k * 32, 16, 1024);
            umma_bf16(d_tmem, ad, bd, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // -------------------------------------------------- epilogue: 16 warps. TMEM lane quadrant = warp % 4 (hardware
    // rule), 64-column chunk of the tile = (warp - 2) / 4. TMEM hands each lane one output ROW. Four warps per scheduler
    // instead of two: the epilogue of the memory-bound 1x1 convolutions is a serial chain per warp (TMEM load -> shift /
    // residual / ReLU -> store) and ncu showed every unit below 60 % with the warps stalled on their own loads; the
    // TMEM loads are software-pipelined 16 columns ahead. Stores go straight from registers (256-bit per lane: L2 merges
    // the sectors and nothing waits on a store). Residual / ReLU-mask tiles are TMA-prefetched into 128B-swizzled
    // shared-memory tiles one output tile ahead and read row-per-lane without bank conflicts; scale / shift vectors are
    // staged in shared memory per n-tile. The "manual" path (nearest-2x upsampled FPN residual, residual + mask
    // together, Cout < 64) uses plain loads.
    const int quad = warp & 3;
    const int ew = warp - 2;                     // 0..15
    const int c0 = (ew >> 2) * 64;               // this warp's 64-column chunk inside the tile
    const int et = threadIdx.x - 64;             // 0..511 among the epilogue threads
    const bool has_chunk = c0 < a.block_n;
    const int cw = min(64, a.block_n - c0);
    const bool use_aux = a.aux_kind != 0 && has_chunk;
    // Memory-bound convolutions finish their MMAs long before the epilogue gets to the tile, so epilogues run back to
    // back: an aux tile requested only after the previous tile was consumed exposes the whole DRAM latency on every tile
    // (ncu: ~30 % of the epilogue warps' samples sat in this mbarrier wait). With aux_dbl each warp owns two staging
    // tiles and requests tile i+1's chunk BEFORE it processes tile i.
    const bool dbl = use_aux && a.aux_dbl;
    // aux_kind 3 (residual AND ReLU mask: the block-output ReLU backward fused into the dgrad that produces the gradient):
    // two tiles per warp, slot 0 = residual, slot 1 = mask, both loads arrive on the warp's first barrier.
    const bool both = a.aux_kind == 3;
    const uint32_t aux_tx = both ? 2 * EPI_TILE_BYTES : EPI_TILE_BYTES;
    uint8_t* atile0 = aux_stage + ew * ((a.aux_dbl || both) ? 2 : 1) * EPI_TILE_BYTES;
    uint64_t* my_aux_bar = aux_bar + ew * 2;
    // out_tma: one 4 KiB staging tile per warp behind the aux tiles; the lane's 8 x 16-byte row goes there (same 128B swizzle
    // as the tensor map) and lane 0 writes the [32 rows x 64 columns] tile with ONE bulk tensor store: full 128-byte lines to
    // L2 and 32 shared-memory wavefronts instead of 128 scattered-sector global-store wavefronts per warp and tile.
    uint8_t* otile_sep = aux_stage + a.aux_bytes + ew * EPI_TILE_BYTES;
    uint32_t acc = 0, acc_phase = 0, aux_phase = 0, it = 0;
    int staged_n_tile = -1;
    if (use_aux && lane == 0 && (int)blockIdx.x < num_tiles) {   // aux tile of this CTA's first tile
      const int t = blockIdx.x;
      const int n_tile = t % a.n_tiles, m_tile = t / a.n_tiles;
      const int lv = level_of(a.lt, m_tile);
      mbar_arrive_expect_tx(&my_aux_bar[0], aux_tx);
      tma_load_2d(atile0, &tmap_aux, &my_aux_bar[0], n_tile * a.block_n + c0,
                  a.lt.row_off[lv] + (m_tile - a.lt.tile_off[lv]) * BM + quad * 32);
      if (both)
        tma_load_2d(atile0 + EPI_TILE_BYTES, &tmap_aux2, &my_aux_bar[0], n_tile * a.block_n + c0,
                    a.lt.row_off[lv] + (m_tile - a.lt.tile_off[lv]) * BM + quad * 32);
    }
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      if (dbl) {                        // request the next tile's chunk into the other slot (last read in tile it-1)
        __syncwarp();
        const int tn = t + gridDim.x;
        if (lane == 0 && a.out_tma == 2) bulk_wait_read0();     // ... and by the in-place store of tile it-1
        if (lane == 0 && tn < num_tiles) {
          const int n_tile2 = tn % a.n_tiles, m_tile2 = tn / a.n_tiles;
          const int lv2 = level_of(a.lt, m_tile2);
          const uint32_t nb = (it + 1) & 1;
          mbar_arrive_expect_tx(&my_aux_bar[nb], EPI_TILE_BYTES);
          tma_load_2d(atile0 + nb * EPI_TILE_BYTES, &tmap_aux, &my_aux_bar[nb], n_tile2 * a.block_n + c0,
                      a.lt.row_off[lv2] + (m_tile2 - a.lt.tile_off[lv2]) * BM + quad * 32);
        }
      }
      uint8_t* atile = atile0 + (dbl ? (it & 1) : 0) * EPI_TILE_BYTES;
      uint8_t* otile = a.out_tma == 2 ? atile : otile_sep;
      const int n_tile = t % a.n_tiles, m_tile = t / a.n_tiles;
      const int lv = level_of(a.lt, m_tile);
      const int ml = (m_tile - a.lt.tile_off[lv]) * BM + quad * 32 + lane;     // row inside the level
      const bool row_ok = ml < a.lt.M[lv];
      const int m = a.lt.row_off[lv] + ml;                                      // row in the level-major buffer
      const int nbase = n_tile * a.block_n;
      if (n_tile != staged_n_tile) {            // (re)stage scale / shift of this n-tile: uniform across the 16 warps
        if (staged_n_tile >= 0) asm volatile("bar.sync 1, 512;" ::: "memory");
        if (et < 256) {
          const int n = nbase + et;
          s_scale[et] = (a.scale && et < a.block_n && n < a.Cout) ? __ldg(a.scale + n) : 1.f;
          s_shift[et] = (a.shift && et < a.block_n && n < a.Cout) ? __ldg(a.shift + n) : 0.f;
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        staged_n_tile = n_tile;
      }
      size_t rrow = (size_t)m;                 // residual row for the manual path
      if (a.manual && a.residual && a.res_up2 && row_ok) {
        const int PQ = a.P * a.Q;
        const int img = m / PQ, rem = m - img * PQ;
        const int p = rem / a.Q, q = rem - p * a.Q;
        rrow = ((size_t)img * (a.P / 2) + (p >> 1)) * (a.Q / 2) + (q >> 1);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (has_chunk) {
        if (a.out_tma == 1) {            // the previous tile's store must have finished READING the staging tile
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
        }
        if (use_aux) {
          if (dbl) mbar_wait(&my_aux_bar[it & 1], (it >> 1) & 1);
          else mbar_wait(&my_aux_bar[0], aux_phase);
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256 + c0;
        uint32_t v[2][16];
        tmem_ld_32x16(taddr, v[0]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          tmem_ld_wait();
          if (j + 1 < 4 && (j + 1) * 16 < cw) tmem_ld_32x16(taddr + (j + 1) * 16, v[(j + 1) & 1]);   // in flight while j is processed
          const uint32_t(&vj)[16] = v[j & 1];
          const int cj = c0 + j * 16;
          const int nj = nbase + cj;
          if (j * 16 < cw && nj < a.Cout && row_ok) {
            float f[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              // the scale vector is only read when there is one (FrozenBN scales are folded into the packed weights, so
              // the product path has none)
              const float4 hv = *reinterpret_cast<const float4*>(s_shift + cj + 4 * i);
              if (a.scale) {
                const float4 sv = *reinterpret_cast<const float4*>(s_scale + cj + 4 * i);
                f[4 * i] = fmaf(__uint_as_float(vj[4 * i]), sv.x, hv.x);
                f[4 * i + 1] = fmaf(__uint_as_float(vj[4 * i + 1]), sv.y, hv.y);
                f[4 * i + 2] = fmaf(__uint_as_float(vj[4 * i + 2]), sv.z, hv.z);
                f[4 * i + 3] = fmaf(__uint_as_float(vj[4 * i + 3]), sv.w, hv.w);
              } else {
                f[4 * i] = __uint_as_float(vj[4 * i]) + hv.x;
                f[4 * i + 1] = __uint_as_float(vj[4 * i + 1]) + hv.y;
                f[4 * i + 2] = __uint_as_float(vj[4 * i + 2]) + hv.z;
                f[4 * i + 3] = __uint_as_float(vj[4 * i + 3]) + hv.w;
              }
            }
            uint4 x0 = make_uint4(0, 0, 0, 0), x1 = x0, y0 = x0, y1 = x0;   // x: residual, y: mask
            int has_res = 0, has_mask = 0;
            if (use_aux) {
              const uint4 t0 = *reinterpret_cast<const uint4*>(atile + swz(lane, 2 * j));
              const uint4 t1 = *reinterpret_cast<const uint4*>(atile + swz(lane, 2 * j + 1));
              if (a.aux_kind == 2) { y0 = t0; y1 = t1; has_mask = 1; } else { x0 = t0; x1 = t1; has_res = 1; }
              if (both) {
                y0 = *reinterpret_cast<const uint4*>(atile + EPI_TILE_BYTES + swz(lane, 2 * j));
                y1 = *reinterpret_cast<const uint4*>(atile + EPI_TILE_BYTES + swz(lane, 2 * j + 1));
                has_mask = 1;
              }
            } else if (a.manual) {
              if (a.residual) {
                const uint4* rp = reinterpret_cast<const uint4*>(a.residual + rrow * a.ldr + nj);
                x0 = __ldg(rp); x1 = __ldg(rp + 1); has_res = 1;
              }
              if (a.relu_mask) {
                const uint4* mp = reinterpret_cast<const uint4*>(a.relu_mask + (size_t)m * a.ldo + nj);
                y0 = __ldg(mp); y1 = __ldg(mp + 1); has_mask = 1;
              }
            }
            if (has_res) {
              const uint32_t xw[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                f[2 * i] += __uint_as_float(xw[i] << 16);
                f[2 * i + 1] += __uint_as_float(xw[i] & 0xFFFF0000u);
              }
            }
            if (a.relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            uint32_t ow[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) ow[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
            if (has_mask) {
              const uint32_t yw[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) ow[i] = mask_bf16x2(ow[i], yw[i]);
            }
            if (a.out_tma) {
              *reinterpret_cast<uint4*>(otile + swz(lane, 2 * j)) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
              *reinterpret_cast<uint4*>(otile + swz(lane, 2 * j + 1)) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
            } else {
              // one 256-bit store per lane (STG.256): half the LSU wavefronts of two 128-bit stores
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(a.out + (size_t)m * a.ldo + nj),
                           "r"(ow[0]), "r"(ow[1]), "r"(ow[2]), "r"(ow[3]), "r"(ow[4]), "r"(ow[5]), "r"(ow[6]), "r"(ow[7])
                           : "memory");
            }
          }
        }
        if (a.out_tma) {                 // rows past M are clipped by the tensor map (single-level launches only)
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_out, otile, nbase + c0, m_tile * BM + quad * 32);
            bulk_commit();
          }
        }
      }
      // accumulators consumed: release the TMEM buffer, then prefetch the aux tile of this CTA's next tile
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (use_aux && !dbl) {
        aux_phase ^= 1;
        __syncwarp();                   // every lane finished reading the aux tile
        const int tn = t + gridDim.x;
        if (lane == 0 && a.out_tma == 2) bulk_wait_read0();     // the in-place store has read the tile
        if (lane == 0 && tn < num_tiles) {
          const int n_tile2 = tn % a.n_tiles, m_tile2 = tn / a.n_tiles;
          const int lv2 = level_of(a.lt, m_tile2);
          mbar_arrive_expect_tx(&my_aux_bar[0], aux_tx);
          tma_load_2d(atile0, &tmap_aux, &my_aux_bar[0], n_tile2 * a.block_n + c0,
                      a.lt.row_off[lv2] + (m_tile2 - a.lt.tile_off[lv2]) * BM + quad * 32);
          if (both)
            tma_load_2d(atile0 + EPI_TILE_BYTES, &tmap_aux2, &my_aux_bar[0], n_tile2 * a.block_n + c0,
                        a.lt.row_off[lv2] + (m_tile2 - a.lt.tile_off[lv2]) * BM + quad * 32);
        }
      }
    }
  }

  if (a.out_tma && warp >= 2 && lane == 0) bulk_wait0();      // all output tiles written before the CTA retires
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ wgrad
constexpr int WG_MAX_STAGES = 8;                // wgrad operand ring depth: as deep as shared memory allows (<= 8)
constexpr int WG_THREADS = 192;                 // warp0 TMA, warp1 MMA, warps2-5 epilogue
// Pixels (GEMM-K) per pipeline stage: a template parameter. One [PIX pix][64 ch] swizzled block is one TMA request; 64
// pixels give the deepest ring (4 stages at block_n = 256) but twice the requests / barrier hand-offs of 128 pixels
// (2 stages), which is what the 3x3 convolutions (operands re-read from L2 nine times) want; 96 sits between.
__host__ __device__ constexpr int wg_blk_bytes(int pix) { return pix * 64 * 2; }
__host__ __device__ constexpr int wg_stage_bytes(int block_n, int pix) { return (2 + block_n / 64) * wg_blk_bytes(pix); }
inline int wg_pick_stages(int block_n, int pix) {
  int s = (SMEM_LIMIT - 1024 - 256) / wg_stage_bytes(block_n, pix);
  return s > WG_MAX_STAGES ? WG_MAX_STAGES : s;
}

struct ConvWgradArgs {
  LevelTable lt;        // tile_off = prefix of WG_PIX-pixel blocks per level
  int Mpix, Cout, Cin, R, S, P, Q, stride, pad;
  int block_n;      // input-channel tile width (64 / 128 / 256)
  int c_tiles, n_tiles, taps;
  int kb_total, kb_per_split;
  int stages;
  const float* scale;   // optional per-Cout factor (FrozenBN scale)
  float* dw;            // [Cout, R*S*Cin] fp32, accumulated
  int cout_store;       // rows >= cout_store are not written (zero-padded fused predictors)
};

template <int WG_PIX>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ TmapSet tmaps_x,
                  const ConvWgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a __shared__ pointer: LDS / STS, not generic LD / ST
  const int STAGES = a.stages;
  constexpr int WG_BLK_BYTES = wg_blk_bytes(WG_PIX);
  constexpr int WG_A_BYTES = 2 * WG_BLK_BYTES;    // 128 output channels
  const int WG_STAGE_BYTES = wg_stage_bytes(a.block_n, WG_PIX);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * WG_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + WG_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + WG_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int tile = blockIdx.x;
  const int c_tile = tile % a.c_tiles; tile /= a.c_tiles;
  const int tap = tile % a.taps;       tile /= a.taps;
  const int n_tile = tile;
  const int kb_begin = blockIdx.y * a.kb_per_split;
  const int kb_end = min(a.kb_total, kb_begin + a.kb_per_split);
  if (kb_begin >= kb_end) return;   // uniform for the whole CTA

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_g);
    prefetch_tmap(&tmaps_x.m[0]);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();
  griddep_launch();
  const uint32_t tmem_base = *tmem_slot;
  const int nblk = a.block_n / 64;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t tx_bytes = (2 + nblk) * WG_BLK_BYTES;
      const int r = tap / a.S, s = tap - r * a.S;
      // incremental pixel walk (no divisions per 64-pixel block: the producer thread must issue 2 + nblk TMA loads
      // every ~500 tensor-pipe cycles)
      int lv = level_of(a.lt, kb_begin);
      const CUtensorMap* tmap_x = &tmaps_x.m[lv];
      int Q = a.lt.Q[lv], P = a.lt.P[lv];
      int pix0 = (kb_begin - a.lt.tile_off[lv]) * WG_PIX;
      int img = pix0 / (P * Q), rem = pix0 - img * (P * Q);
      int p0 = rem / Q, q0 = rem - p0 * Q;
      int next_lv_kb = a.lt.tile_off[lv + 1];
      int grow = a.lt.row_off[lv] + pix0;
      const int c_base = c_tile * a.block_n, ncol = n_tile * 128;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        if (kb == next_lv_kb) {                      // first block of the next pyramid level
          ++lv;
          tmap_x = &tmaps_x.m[lv];
          Q = a.lt.Q[lv]; P = a.lt.P[lv];
          img = 0; p0 = 0; q0 = 0;
          next_lv_kb = a.lt.tile_off[lv + 1];
          grow = a.lt.row_off[lv];
        }
        const int w0 = q0 * a.stride - a.pad, h0 = p0 * a.stride - a.pad;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * WG_STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
        for (int j = 0; j < 2; ++j)
          tma_load_2d(sa + j * WG_BLK_BYTES, &tmap_g, &full_bar[stage], ncol + j * 64,
                      grow);     // dY rows past the level's end meet zero-filled (OOB) X rows
        for (int j = 0; j < nblk; ++j)
          tma_load_im2col_4d(sa + WG_A_BYTES + j * WG_BLK_BYTES, tmap_x, &full_bar[stage],
                             c_base + j * 64, w0, h0, img, (uint16_t)s, (uint16_t)r);
        if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1; }
        grow += WG_PIX;
        q0 += WG_PIX;
        while (q0 >= Q) { q0 -= Q; ++p0; }
        while (p0 >= P) { p0 -= P; ++img; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, a.block_n, 1, 1);
      uint32_t stage = 0, phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * WG_STAGE_BYTES);
        const uint32_t sb = sa + WG_A_BYTES;
#pragma unroll
        for (int k = 0; k < WG_PIX / 16; ++k) {
          // MN-major SW128: 64-channel blocks LBO apart, 8-pixel groups SBO apart.
          const uint64_t ad = umma_smem_desc_sw128(sa + k * 2048, WG_BLK_BYTES, 1024);
          const uint64_t bd = umma_smem_desc_sw128(sb + k * 2048, WG_BLK_BYTES, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (kb != kb_begin) || (k != 0));
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull_bar);
    }
  } else {
    const int quad = warp & 3;
    const int n = n_tile * 128 + quad * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const size_t K = (size_t)a.R * a.S * a.Cin;
    const float sc = (a.scale && n < a.cout_store) ? __ldg(a.scale + n) : 1.f;
    for (int c = 0; c < a.block_n; c += 16) {
      uint32_t v[16];
      tmem_ld_32x16(taddr + c, v);
      tmem_ld_wait();
      if (n < a.cout_store) {
        // 128-bit vector reductions (RED.128): a quarter of the L2 atomic transactions of scalar adds — with split-K
        // every CTA flushes a 128 x block_n fp32 tile, which is a visible tail for the shorter reductions
        float* dst = a.dw + (size_t)n * K + (size_t)tap * a.Cin + c_tile * a.block_n + c;
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          atomicAdd(reinterpret_cast<float4*>(dst + i),
                    make_float4(__uint_as_float(v[i]) * sc, __uint_as_float(v[i + 1]) * sc, __uint_as_float(v[i + 2]) * sc,
                                __uint_as_float(v[i + 3]) * sc));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------ probe
// Debug: one im2col TMA load -> raw smem dump (used by tests to pin the descriptor semantics).
__global__ void im2col_probe_kernel(const __grid_constant__ CUtensorMap tmap, int c, int w, int h,
                                    int n, int off_w, int off_h, int bytes, uint8_t* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a __shared__ pointer: LDS / STS, not generic LD / ST
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, bytes);
    tma_load_im2col_4d(smem, &tmap, &bar, c, w, h, n, (uint16_t)off_w, (uint16_t)off_h);
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return ut2_sm_budget(g_num_sms);
}

static int pick_block_n(int cout) {
  if (cout % 16) return -1;
  if (cout <= 256) return cout;
  if (cout % 256 == 0) return 256;
  if (cout % 128 == 0) return 128;
  return -1;
}

}  // namespace ut2

using namespace ut2;

// Shared launcher. num_levels == 1: an ordinary convolution on [N, H, W, Cin]. num_levels > 1: one launch over a
// level-major pyramid (levels [N, H_l, W_l, Cin] laid back to back in x, outputs laid back to back in y; stride 1).
static int conv_fwd_launch(const void* x, int num_levels, const int* hw, int N, int Cin, const void* w, int Cout, int R,
                           int S, int stride, int pad, const float* scale, const float* shift, const void* residual,
                           int res_up2, const void* relu_mask, int relu, void* y, void* stream) {
  if (!x || !w || !y) return ut2_fail(-1, "conv_fwd: null pointer");
  if (Cin % 8 || Cin <= 0) return ut2_fail(-2, "conv_fwd: Cin must be a multiple of 8");
  int block_n = pick_block_n(Cout);
  if (block_n < 0) return ut2_fail(-3, "conv_fwd: unsupported Cout (need %16==0; >256 needs %128==0)");
  if (stride < 1 || stride > 8 || pad < 0 || R < 1 || S < 1) return ut2_fail(-4, "conv_fwd: bad geometry");
  if (num_levels < 1 || num_levels > MAX_LV) return ut2_fail(-4, "conv_fwd: 1..5 levels");
  if (num_levels > 1 && (stride != 1 || res_up2)) return ut2_fail(-4, "conv_fwd: multi-level launches are stride 1, no res_up2");
  if (R == 3 && S == 3 && stride == 1 && pad == 1 && !scale && !residual) {
    const int rc = conv3x3_halo_try(x, num_levels, hw, N, Cin, w, Cout, shift, relu_mask, relu, y, num_sms(), stream);
    if (rc) return rc < 0 ? rc : 0;      // taken (or failed): narrow 3x3 convolutions run on 2-D patches, see conv3x3_halo.cu
  }
  ConvFwdArgs a;
  TmapSet tx;
  a.lt.num = num_levels;
  a.lt.tile_off[0] = 0;
  long long in_rows = 0, out_rows = 0;
  for (int l = 0; l < MAX_LV; ++l) {
    const int ll = l < num_levels ? l : num_levels - 1;
    const int H = hw[2 * ll], W = hw[2 * ll + 1];
    const int P = (H + 2 * pad - R) / stride + 1, Q = (W + 2 * pad - S) / stride + 1;
    if (P <= 0 || Q <= 0) return ut2_fail(-4, "conv_fwd: empty output");
    if (l < num_levels) {
      a.lt.P[l] = P; a.lt.Q[l] = Q; a.lt.M[l] = N * P * Q;
      a.lt.row_off[l] = (int)out_rows;
      a.lt.tile_off[l + 1] = a.lt.tile_off[l] + (a.lt.M[l] + BM - 1) / BM;
      int rc = make_tmap_im2col_bf16(&tx.m[l], static_cast<const __nv_bfloat16*>(x) + in_rows * Cin, N, H, W, Cin, R, S,
                                     stride, pad, 64, BM);
      if (rc) return ut2_fail(rc, "conv_fwd: activation tensor map encode failed");
      in_rows += (long long)N * H * W;
      out_rows += a.lt.M[l];
    } else {
      a.lt.P[l] = a.lt.P[ll]; a.lt.Q[l] = a.lt.Q[ll]; a.lt.M[l] = 0; a.lt.row_off[l] = 0;
      if (l + 1 <= MAX_LV) a.lt.tile_off[l + 1] = a.lt.tile_off[l];
      tx.m[l] = tx.m[ll];
    }
  }
  // Under-filled grids (small batches, the deep trunk / top pyramid levels: fewer tiles than half the SMs): narrower tiles
  // put more SMs to work; the A tiles they re-read come from L2. UT2_NARROW=0 disables (A/B runs).
  {
    static int narrow_on = -1;
    if (narrow_on < 0) { const char* e = getenv("UT2_NARROW"); narrow_on = e ? atoi(e) : 1; }
    const int m_tiles = a.lt.tile_off[num_levels];
    while (narrow_on && block_n >= 128 && block_n % 32 == 0 && Cout % (block_n / 2) == 0 &&
           m_tiles * ((Cout + block_n - 1) / block_n) * 2 <= num_sms())
      block_n /= 2;
  }
  // residual + mask (aux_kind 3): the 128 KiB of aux tiles leave two operand stages next to a 256-wide tile. Measured against
  // 128-wide tiles (64 KiB of aux tiles in use, five stages; UT2_AUX3_BN=128) on the res3 / res4 / res5 conv1 data-gradients,
  // L2 flushed: 161 / 94 / 70 us vs 217 / 121 / 72 us — these launches are bound by their epilogue, not by the main loop.
  if (residual && relu_mask && !res_up2 && Cout >= 256 && Cout % 128 == 0 && block_n > 128) {
    static int bn3 = -1;
    if (bn3 < 0) { const char* e = getenv("UT2_AUX3_BN"); bn3 = e ? atoi(e) : 256; }
    if (bn3 == 128) block_n = 128;
  }
  const int P = a.lt.P[0], Q = a.lt.Q[0];
  a.M = (int)out_rows; a.Cout = Cout; a.ldo = Cout;
  a.block_n = block_n; a.n_tiles = (Cout + block_n - 1) / block_n; a.m_tiles = a.lt.tile_off[num_levels];
  a.P = P; a.Q = Q; a.stride = stride; a.pad = pad; a.R = R; a.S = S; a.Cin = Cin;
  a.relu = relu; a.scale = scale; a.shift = shift;
  a.residual = static_cast<const __nv_bfloat16*>(residual); a.ldr = Cout;
  a.res_up2 = res_up2; a.relu_mask = static_cast<const __nv_bfloat16*>(relu_mask);
  if (res_up2 && ((P & 1) || (Q & 1))) return ut2_fail(-4, "conv_fwd: res_up2 needs even output H, W");
  a.out = static_cast<__nv_bfloat16*>(y);
  a.manual = (Cout < 64 && (residual || relu_mask)) || (residual && res_up2);
  a.aux_kind = a.manual ? 0 : ((residual && relu_mask) ? 3 : (residual ? 1 : (relu_mask ? 2 : 0)));
  // double-buffered aux tiles for the most epilogue-bound convolutions (K <= 128: res2 / res3 conv3 + shortcut, +11 %
  // measured); from K = 256 on the two operand stages that are left cost more than the exposed aux latency (K = 512:
  // -25 %), and the tensor-bound ones hide it behind their main loop anyway. UT2_AUX_DBL=0 disables (A/B runs).
  static int dbl_on = -1;
  if (dbl_on < 0) { const char* e = getenv("UT2_AUX_DBL"); dbl_on = e ? atoi(e) : 1; }
  a.aux_dbl = (dbl_on && a.aux_kind != 0 && a.aux_kind != 3 && R * S * Cin <= 128) ? 1 : 0;
  // kind 3 keeps two tiles per ACTIVE epilogue warp (four warps per 64-column chunk of the tile)
  const int aux_tiles = a.aux_kind == 3 ? 2 * 4 * ((block_n + 63) / 64) : (a.aux_kind ? (a.aux_dbl ? 32 : 16) : 0);
  // TMA-store epilogue for the memory-bound 1x1 convolutions (UT2_TMA_STORE=0 disables, =2 forces it for every single-level
  // launch): 16 more staging tiles; not next to the 32 tiles of aux_dbl / aux_kind 3 with 256-wide tiles (no room).
  static int tma_store = -1;
  if (tma_store < 0) { const char* e = getenv("UT2_TMA_STORE"); tma_store = e ? atoi(e) : 1; }
  a.out_tma = 0;
  if (tma_store && num_levels == 1 && !a.manual && Cout % 64 == 0 && (tma_store == 2 || R * S == 1)) {
    const int num_kb = R * S * ((Cin + BK - 1) / BK);
    const int want = num_kb < 3 ? num_kb : 3;
    if (a.aux_kind == 0)
      a.out_tma = (tma_store == 2 || num_kb <= 4) && pick_stages(block_n, 16, 1024) >= (want < 2 ? 2 : want) ? 1 : 0;
    else if (a.aux_dbl || tma_store == 2)   // the lane's aux row has been consumed when its output row is ready: stage in place
      a.out_tma = 2;                        // (single aux tile: the next prefetch would wait for the store -> measured slower)
  }
  a.aux_bytes = aux_tiles * EPI_TILE_BYTES;
  const int aux_tiles_total = aux_tiles + (a.out_tma == 1 ? 16 : 0);
  a.no_pad = ((a.aux_dbl || a.aux_kind == 3) && pick_stages(block_n, aux_tiles_total, 1024) < 2) ? 1 : 0;
  const int smem_pad = a.no_pad ? 0 : 1024;
  a.stages = pick_stages(block_n, aux_tiles_total, smem_pad);
  if (a.stages < 2) return ut2_fail(-5, "conv_fwd: shared memory budget");
  CUtensorMap tw, to, ta;
  int rc = make_tmap_2d_bf16(&tw, w, Cout, (uint64_t)R * S * Cin, (uint64_t)R * S * Cin, 64, block_n);
  if (rc) return ut2_fail(rc, "conv_fwd: weight tensor map encode failed");
  const uint32_t bc = Cout < 64 ? Cout : 64;
  rc = make_tmap_2d_bf16(&to, y, a.M, Cout, Cout, bc, 32);
  if (rc) return ut2_fail(rc, "conv_fwd: output tensor map encode failed");
  ta = to;
  CUtensorMap ta2 = to;
  if (a.aux_kind) {
    rc = make_tmap_2d_bf16(&ta, a.aux_kind == 2 ? relu_mask : residual, a.M, Cout, Cout, bc, 32);
    if (rc) return ut2_fail(rc, "conv_fwd: aux tensor map encode failed");
  }
  if (a.aux_kind == 3) {
    rc = make_tmap_2d_bf16(&ta2, relu_mask, a.M, Cout, Cout, bc, 32);
    if (rc) return ut2_fail(rc, "conv_fwd: mask tensor map encode failed");
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return ut2_fail((int)e, "conv_fwd: cudaFuncSetAttribute");
    attr_set = true;
  }
  const int tiles = a.m_tiles * a.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  ut2_launch_pdl(conv_fwd_kernel, dim3(grid), dim3(NUM_THREADS), smem_bytes(a.stages, block_n, aux_tiles_total, smem_pad),
                 static_cast<cudaStream_t>(stream),
                 ut2_est_us(2.0 * a.M * Cout * R * S * Cin, 2.0 * a.M * (Cin + Cout * (1.0 + (a.aux_kind != 0) + (a.aux_kind == 3)))),
                 tx, tw, to, ta, ta2, a);
  return ut2_check_launch("conv_fwd");
}

extern "C" int ut2_conv2d_nhwc_bf16_fwd(const void* x, int N, int H, int W, int Cin, const void* w,
                                        int Cout, int R, int S, int stride, int pad,
                                        const float* scale, const float* shift,
                                        const void* residual, int res_up2, const void* relu_mask, int relu,
                                        void* y, void* stream) {
  const int hw[2] = {H, W};
  return conv_fwd_launch(x, 1, hw, N, Cin, w, Cout, R, S, stride, pad, scale, shift, residual, res_up2, relu_mask, relu, y,
                         stream);
}

// One launch over a level-major pyramid: x = levels [N, H_l, W_l, Cin] back to back, y likewise with Cout channels
// (stride 1). hw is a HOST array [H_0, W_0, H_1, W_1, ...]. The shared-weight FCOS tower / predictor convolutions
// (fcos/fcos.py:338-376 applies the same modules to all five FPN levels) run as one grid instead of five.
extern "C" int ut2_conv2d_levels_bf16_fwd(const void* x, int num_levels, const int* hw, int N, int Cin, const void* w,
                                          int Cout, int R, int S, int pad, const float* scale, const float* shift,
                                          const void* residual, const void* relu_mask, int relu, void* y, void* stream) {
  return conv_fwd_launch(x, num_levels, hw, N, Cin, w, Cout, R, S, 1, pad, scale, shift, residual, 0, relu_mask, relu, y,
                         stream);
}

static int conv_wgrad_launch(const void* x, int num_levels, const int* hw, int N, int Cin, const void* dy, int Cout, int R,
                             int S, int stride, int pad, const float* scale, float* dw, int cout_store, void* stream) {
  if (!x || !dy || !dw) return ut2_fail(-1, "conv_wgrad: null pointer");
  if (Cin % 64 || Cin <= 0) return ut2_fail(-2, "conv_wgrad: Cin must be a multiple of 64");
  if (Cout % 8) return ut2_fail(-3, "conv_wgrad: Cout must be a multiple of 8");
  if (num_levels < 1 || num_levels > MAX_LV) return ut2_fail(-4, "conv_wgrad: 1..5 levels");
  if (num_levels > 1 && stride != 1) return ut2_fail(-4, "conv_wgrad: multi-level launches are stride 1");
  ConvWgradArgs a;
  TmapSet tx;
  // Stage size (measured over the R50-FPN / head shapes, tools/bench_conv.py): 256-wide input-channel tiles want 96
  // pixels (3 stages of 72 KiB) unless the reduction is short and spread over few output tiles, everything narrower
  // wants 128 pixels. UT2_WG_PIX=64|96|128 overrides for experiments.
  static int pix_override = -1;
  if (pix_override < 0) {
    const char* e = getenv("UT2_WG_PIX");
    pix_override = e ? atoi(e) : 0;
    if (pix_override != 64 && pix_override != 96 && pix_override != 128) pix_override = 0;
  }
  a.lt.num = num_levels;
  a.lt.tile_off[0] = 0;
  long long in_rows = 0, out_rows = 0;
  a.block_n = Cin % 256 == 0 ? 256 : (Cin % 128 == 0 ? 128 : 64);
  a.c_tiles = Cin / a.block_n; a.n_tiles = (Cout + 127) / 128; a.taps = R * S;
  const int out_tiles = a.c_tiles * a.n_tiles * a.taps;
  long long m_total = 0;
  for (int l = 0; l < num_levels; ++l)
    m_total += (long long)N * ((hw[2 * l] + 2 * pad - R) / stride + 1) * ((hw[2 * l + 1] + 2 * pad - S) / stride + 1);
  int WG_PIX = a.block_n == 256 ? ((m_total < 32768 && out_tiles < 32) ? 128 : 96) : 128;
  if (pix_override) WG_PIX = pix_override;
  for (int l = 0; l < MAX_LV; ++l) {
    const int ll = l < num_levels ? l : num_levels - 1;
    const int H = hw[2 * ll], W = hw[2 * ll + 1];
    const int P = (H + 2 * pad - R) / stride + 1, Q = (W + 2 * pad - S) / stride + 1;
    if (P <= 0 || Q <= 0) return ut2_fail(-4, "conv_wgrad: empty output");
    if (l < num_levels) {
      a.lt.P[l] = P; a.lt.Q[l] = Q; a.lt.M[l] = N * P * Q;
      a.lt.row_off[l] = (int)out_rows;
      a.lt.tile_off[l + 1] = a.lt.tile_off[l] + (a.lt.M[l] + WG_PIX - 1) / WG_PIX;
      int rc = make_tmap_im2col_bf16(&tx.m[l], static_cast<const __nv_bfloat16*>(x) + in_rows * Cin, N, H, W, Cin, R, S,
                                     stride, pad, 64, WG_PIX);
      if (rc) return ut2_fail(rc, "conv_wgrad: activation tensor map encode failed");
      in_rows += (long long)N * H * W;
      out_rows += a.lt.M[l];
    } else {
      a.lt.P[l] = a.lt.P[ll]; a.lt.Q[l] = a.lt.Q[ll]; a.lt.M[l] = 0; a.lt.row_off[l] = 0;
      if (l + 1 <= MAX_LV) a.lt.tile_off[l + 1] = a.lt.tile_off[l];
      tx.m[l] = tx.m[ll];
    }
  }
  a.Mpix = (int)out_rows; a.Cout = Cout; a.Cin = Cin; a.R = R; a.S = S; a.P = a.lt.P[0]; a.Q = a.lt.Q[0];
  a.stride = stride; a.pad = pad;
  a.kb_total = a.lt.tile_off[num_levels];
  // split-K so that the grid is (at most) one full wave of CTAs, but keep >= 32 pixel blocks per CTA: the fp32
  // atomic epilogue (128 x block_n values per CTA) must stay small next to the main loop
  int splits = num_sms() / out_tiles;
  if (splits > a.kb_total / (2048 / WG_PIX)) splits = a.kb_total / (2048 / WG_PIX);
  if (splits < 1) splits = 1;
  a.kb_per_split = (a.kb_total + splits - 1) / splits;
  splits = (a.kb_total + a.kb_per_split - 1) / a.kb_per_split;
  a.scale = scale; a.dw = dw;
  a.stages = wg_pick_stages(a.block_n, WG_PIX);
  const int wg_smem = a.stages * wg_stage_bytes(a.block_n, WG_PIX) + 1024 + 256;
  a.cout_store = (cout_store > 0 && cout_store < Cout) ? cout_store : Cout;
  CUtensorMap tg;
  int rc = make_tmap_2d_bf16(&tg, dy, a.Mpix, Cout, Cout, 64, WG_PIX);
  if (rc) return ut2_fail(rc, "conv_wgrad: dY tensor map encode failed");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_wgrad_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_wgrad_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return ut2_fail((int)e, "conv_wgrad: cudaFuncSetAttribute");
    attr_set = true;
  }
  dim3 grid(out_tiles, splits);
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double est = ut2_est_us(2.0 * a.Mpix * a.Cout * a.R * a.S * a.Cin, 2.0 * a.Mpix * (a.Cin + a.Cout));
  if (WG_PIX == 128)
    ut2_launch_pdl(conv_wgrad_kernel<128>, grid, dim3(WG_THREADS), wg_smem, st, est, tg, tx, a);
  else if (WG_PIX == 96)
    ut2_launch_pdl(conv_wgrad_kernel<96>, grid, dim3(WG_THREADS), wg_smem, st, est, tg, tx, a);
  else
    ut2_launch_pdl(conv_wgrad_kernel<64>, grid, dim3(WG_THREADS), wg_smem, st, est, tg, tx, a);
  return ut2_check_launch("conv_wgrad");
}

extern "C" int ut2_conv2d_nhwc_bf16_wgrad(const void* x, int N, int H, int W, int Cin,
                                          const void* dy, int Cout, int R, int S, int stride,
                                          int pad, const float* scale, float* dw, int cout_store,
                                          void* stream) {
  const int hw[2] = {H, W};
  return conv_wgrad_launch(x, 1, hw, N, Cin, dy, Cout, R, S, stride, pad, scale, dw, cout_store, stream);
}

// Weight gradient of a shared-weight convolution over a level-major pyramid in one launch (the reduction simply runs
// over the pixels of all levels). Same layout contract as ut2_conv2d_levels_bf16_fwd.
extern "C" int ut2_conv2d_levels_bf16_wgrad(const void* x, int num_levels, const int* hw, int N, int Cin, const void* dy,
                                            int Cout, int R, int S, int pad, const float* scale, float* dw, int cout_store,
                                            void* stream) {
  return conv_wgrad_launch(x, num_levels, hw, N, Cin, dy, Cout, R, S, 1, pad, scale, dw, cout_store, stream);
}

extern "C" int ut2_debug_im2col_probe(const void* x, int N, int H, int W, int C, int R, int S,
                                      int stride, int pad, int pixels, int c, int w, int h, int n,
                                      int off_w, int off_h, void* out, void* stream) {
  CUtensorMap tx;
  int rc = make_tmap_im2col_bf16(&tx, x, N, H, W, C, R, S, stride, pad, 64, pixels);
  if (rc) return ut2_fail(rc, "probe: tensor map encode failed");
  const int bytes = pixels * 128;
  cudaFuncSetAttribute(im2col_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 1024);
  im2col_probe_kernel<<<1, 128, bytes + 1024, static_cast<cudaStream_t>(stream)>>>(
      tx, c, w, h, n, off_w, off_h, bytes, static_cast<uint8_t*>(out));
  return ut2_check_launch("im2col_probe");
}
