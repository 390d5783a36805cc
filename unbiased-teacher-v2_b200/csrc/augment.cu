// Strong augmentation of the two-crop pipeline on the device: uint8 CHW in, uint8 CHW out.
//
// Reference: ubteacher/data/detection_utils.py:8-46 (build_strong_augmentation) and
// ubteacher/data/transforms/augmentation_impl.py:7-23 (GaussianBlur), i.e. torchvision transforms on PIL images:
//   RandomApply([ColorJitter(.4,.4,.4,.1)], .8) -> RandomGrayscale(.2) -> RandomApply([GaussianBlur([.1, 2.])], .5)
//   -> ToTensor -> RandomErasing x3 (value="random") -> ToPILImage.
// The arithmetic restates Pillow's integer / float algorithms bit for bit (oracle/ut2_aug_oracle.py is pinned against
// Pillow + torchvision): Image.blend in float32 with truncation, fixed-point ITU-R 601 luma, float32/float64 HSV round
// trip with truncating casts, the 24-bit fixed-point extended box blur (3 horizontal + 3 vertical passes, uint8
// between passes), Tensor.byte() wrap-around of the erase noise. All random draws are made on the host (same order as
// torchvision) and arrive in a per-image parameter table; only the erase noise may be generated here (counter hash ->
// Box-Muller), or passed explicitly for parity tests. The uint8 round trip ToTensor -> ToPILImage is the identity.
//
// Every kernel is a flat coalesced pass over planar uint8 (HBM-bound: <= 6 B/pixel per pass); one launch covers the
// batch: grid = (blocks per image, N), images that do not take part in a pass exit at once.
#include "ut2_internal.h"
#include <math.h>
#include <stdint.h>

namespace {

struct AugImage {            // 168 bytes, mirrored by ubteacher/data/gpu_augmentation.py (struct format in the header)
  const uint8_t* src;        // uint8 [3, h, w]
  uint8_t* dst;              // uint8 [3, h, w]
  uint8_t* tmp;              // uint8 [3, h, w] scratch (blur ping-pong); may be NULL when blur_r < 0
  const float* noise[3];     // optional float32 [3, eh, ew] per erase (NULL: hashed N(0,1))
  int h, w;
  int order[4];              // ColorJitter op of slot k: 0 brightness, 1 contrast, 2 saturation, 3 hue; -1 = none
  float factor[4];           // brightness, contrast, saturation, hue factors
  int hue_shift;             // uint8(int32(hue_factor * 255))
  int gray;
  int blur_r;                // box radius int part, < 0 = no blur
  unsigned int blur_ww, blur_fw;
  int n_erase;
  int ei[3], ej[3], eh[3], ew[3];
  unsigned int seed;
  int pad_;
};
static_assert(sizeof(AugImage) == 168, "AugImage layout is part of the ABI");

__device__ __forceinline__ unsigned int luma(unsigned int r, unsigned int g, unsigned int b) {
  return (r * 19595u + g * 38470u + b * 7471u + 0x8000u) >> 16;
}
// Image.blend(deg, img, alpha): float32 in1 + alpha * (in2 - in1); truncation, clipping only when extrapolating
__device__ __forceinline__ unsigned int blend1(int deg, int x, float alpha, bool interp) {
  const float t = __fadd_rn((float)deg, __fmul_rn(alpha, (float)(x - deg)));
  if (interp) return (unsigned int)(int)t & 255u;
  return t <= 0.f ? 0u : (t >= 255.f ? 255u : (unsigned int)(int)t);
}
__device__ __forceinline__ unsigned int clip8(int v) { return v < 0 ? 0u : (v > 255 ? 255u : (unsigned int)v); }

// libImaging/Convert.c rgb2hsv_row -> uint8 shift of H -> hsv2rgb_row
__device__ __forceinline__ void hue_px(unsigned int& r, unsigned int& g, unsigned int& b, unsigned int shift) {
  const unsigned int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
  unsigned int uh = 0, us = 0;
  const unsigned int uv = maxc;
  if (minc != maxc) {
    const float cr = (float)(maxc - minc);
    const float s = __fdiv_rn(cr, (float)maxc);
    const float rc = __fdiv_rn((float)(maxc - r), cr), gc = __fdiv_rn((float)(maxc - g), cr), bc = __fdiv_rn((float)(maxc - b), cr);
    float h;
    if (r == maxc) h = __fsub_rn(bc, gc);
    else if (g == maxc) h = (float)(2.0 + (double)rc - (double)bc);
    else h = (float)(4.0 + (double)gc - (double)rc);
    h = (float)fmod((double)h / 6.0 + 1.0, 1.0);
    uh = clip8((int)((double)h * 255.0));
    us = clip8((int)((double)s * 255.0));
  }
  uh = (uh + shift) & 255u;
  if (us == 0) { r = g = b = uv; return; }
  const double hf = (double)uh * 6.0 / 255.0;
  const double fi = floor(hf);
  const double f = (double)(float)(hf - fi);
  const double fs = (double)__fdiv_rn((float)us, 255.f);
  const double v = (double)uv;
  const unsigned int p = clip8((int)rint(v * (1.0 - fs)));
  const unsigned int q = clip8((int)rint(v * (1.0 - fs * f)));
  const unsigned int t = clip8((int)rint(v * (1.0 - fs * (1.0 - f))));
  switch ((int)fi % 6) {
    case 0: r = uv; g = t; b = p; break;
    case 1: r = q; g = uv; b = p; break;
    case 2: r = p; g = uv; b = t; break;
    case 3: r = p; g = q; b = uv; break;
    case 4: r = t; g = p; b = uv; break;
    default: r = uv; g = p; b = q; break;
  }
}

__global__ void __launch_bounds__(256) aug_copy_kernel(const AugImage* __restrict__ tab) {
  const AugImage& im = tab[blockIdx.y];
  const size_t n = (size_t)3 * im.h * im.w;
  if ((n & 15) == 0 && ((reinterpret_cast<uintptr_t>(im.dst) | reinterpret_cast<uintptr_t>(im.src)) & 15) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(im.src);
    uint4* d4 = reinterpret_cast<uint4*>(im.dst);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (n >> 4); i += (size_t)gridDim.x * blockDim.x) d4[i] = s4[i];
    return;
  }
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) im.dst[i] = im.src[i];
}

// sum of the luma over the image, for the images whose ColorJitter slot `slot` is the contrast op
__global__ void __launch_bounds__(256) aug_lsum_kernel(const AugImage* __restrict__ tab, int slot, unsigned long long* __restrict__ lsum) {
  const AugImage& im = tab[blockIdx.y];
  if (im.order[slot] != 1) return;
  const size_t hw = (size_t)im.h * im.w;
  unsigned long long s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x)
    s += luma(im.dst[i], im.dst[hw + i], im.dst[2 * hw + i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(lsum + blockIdx.y * 4 + slot, s);
}

__global__ void __launch_bounds__(256) aug_color_kernel(const AugImage* __restrict__ tab, int slot, const unsigned long long* __restrict__ lsum) {
  const AugImage& im = tab[blockIdx.y];
  const int op = im.order[slot];
  if (op < 0) return;
  const size_t hw = (size_t)im.h * im.w;
  const float f = im.factor[op];
  const bool interp = f >= 0.f && f <= 1.f;
  int mean = 0;
  if (op == 1) mean = (int)((double)lsum[blockIdx.y * 4 + slot] / (double)hw + 0.5);      // ImageStat mean, int(m + 0.5)
  uint8_t* c0 = im.dst; uint8_t* c1 = im.dst + hw; uint8_t* c2 = im.dst + 2 * hw;
  auto px = [&](unsigned int& r, unsigned int& g, unsigned int& b) {
    if (op == 3) {
      hue_px(r, g, b, (unsigned int)im.hue_shift);
    } else {
      const int d = op == 0 ? 0 : (op == 1 ? mean : (int)luma(r, g, b));
      r = blend1(d, (int)r, f, interp); g = blend1(d, (int)g, f, interp); b = blend1(d, (int)b, f, interp);
    }
  };
  if ((hw & 3) == 0 && (reinterpret_cast<uintptr_t>(im.dst) & 3) == 0) {      // 4 pixels per thread: 32-bit plane accesses
    const size_t n4 = hw >> 2;
    uint32_t* p0 = reinterpret_cast<uint32_t*>(c0); uint32_t* p1 = reinterpret_cast<uint32_t*>(c1); uint32_t* p2 = reinterpret_cast<uint32_t*>(c2);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
      const uint32_t wr = p0[i], wg = p1[i], wb = p2[i];
      uint32_t orr = 0, og = 0, ob = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unsigned int r = (wr >> (8 * k)) & 255u, g = (wg >> (8 * k)) & 255u, b = (wb >> (8 * k)) & 255u;
        px(r, g, b);
        orr |= r << (8 * k); og |= g << (8 * k); ob |= b << (8 * k);
      }
      p0[i] = orr; p1[i] = og; p2[i] = ob;
    }
    return;
  }
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
    unsigned int r = c0[i], g = c1[i], b = c2[i];
    px(r, g, b);
    c0[i] = (uint8_t)r; c1[i] = (uint8_t)g; c2[i] = (uint8_t)b;
  }
}

__global__ void __launch_bounds__(256) aug_gray_kernel(const AugImage* __restrict__ tab) {
  const AugImage& im = tab[blockIdx.y];
  if (!im.gray) return;
  const size_t hw = (size_t)im.h * im.w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
    const uint8_t l = (uint8_t)luma(im.dst[i], im.dst[hw + i], im.dst[2 * hw + i]);
    im.dst[i] = l; im.dst[hw + i] = l; im.dst[2 * hw + i] = l;
  }
}

constexpr int BOX_FUSED_R_FWD = 2;      // == BOX_FUSED_R (defined with the fused kernels below)

// One extended-box-blur pass (libImaging/BoxBlur.c): out = (ww * sum_{|k|<=r} in[x+k] + fw * (in[x-r-1] + in[x+r+1]) +
// 2^23) >> 24 with replicated edges, along x (vertical = 0) or y (vertical = 1). pass parity picks the ping-pong side.
__global__ void __launch_bounds__(256) aug_box_kernel(const AugImage* __restrict__ tab, int pass, int vertical) {
  const AugImage& im = tab[blockIdx.y];
  if (im.blur_r <= BOX_FUSED_R_FWD) return;    // radii <= 2 (every radius the reference draws) take the fused kernels below
  const uint8_t* in = (pass & 1) ? im.tmp : im.dst;
  uint8_t* out = (pass & 1) ? im.dst : im.tmp;
  const int h = im.h, w = im.w, r = im.blur_r;
  const size_t hw = (size_t)h * w, n = 3 * hw;
  const unsigned long long ww = im.blur_ww, fw = im.blur_fw;
  const int len = vertical ? h : w, stride = vertical ? w : 1;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i % hw;
    const int y = (int)(pix / w), x = (int)(pix - (size_t)y * w);
    const int pos = vertical ? y : x;
    const uint8_t* line = in + (i - (size_t)pos * stride);
    unsigned long long acc = 0;
    for (int k = -r; k <= r; ++k) acc += line[(size_t)min(max(pos + k, 0), len - 1) * stride];
    const unsigned long long far = (unsigned long long)line[(size_t)max(pos - r - 1, 0) * stride] + line[(size_t)min(pos + r + 1, len - 1) * stride];
    out[i] = (uint8_t)((acc * ww + far * fw + (1ull << 23)) >> 24);
  }
}

// The three passes of one direction in ONE kernel (box radius <= 2, i.e. every Gaussian radius up to ~3.5): a row segment /
// column strip plus a halo of 3 * (r + 1) pixels is staged in shared memory, blurred three times there (uint8 between the
// passes, like Pillow) and written once: 2 instead of 6 bytes moved per pixel, channel and direction. Pillow replicates
// the edge pixel of the CURRENT pass, so after every pass the out-of-image part of the halo is refilled from the freshly
// blurred border pixel.
constexpr int BOX_FUSED_R = 2;
constexpr int BXH_SEG = 224;                    // output pixels per row segment (224 + 2 * 9 <= 256 threads)
constexpr int BXV_TH = 32, BXV_TW = 64;         // rows x columns per strip tile

__device__ __forceinline__ uint8_t box_tap(const uint8_t* b, int t, int stride, int r, unsigned long long ww, unsigned long long fw) {
  unsigned long long acc = 0;
  for (int k = -r; k <= r; ++k) acc += b[(t + k) * stride];
  const unsigned long long far = (unsigned long long)b[(t - r - 1) * stride] + b[(t + r + 1) * stride];
  return (uint8_t)((acc * ww + far * fw + (1ull << 23)) >> 24);
}

// horizontal: dst -> tmp
__global__ void __launch_bounds__(256) aug_box3_h_kernel(const AugImage* __restrict__ tab) {
  const AugImage& im = tab[blockIdx.y];
  if (im.blur_r < 0 || im.blur_r > BOX_FUSED_R) return;
  __shared__ uint8_t buf[2][256];
  const int r = im.blur_r, halo = 3 * (r + 1), L = BXH_SEG + 2 * halo, w = im.w;
  const unsigned long long ww = im.blur_ww, fw = im.blur_fw;
  const int nseg = (w + BXH_SEG - 1) / BXH_SEG;
  const long long items = 3ll * im.h * nseg;
  const int t = threadIdx.x;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const long long row = item / nseg;
    const int x0 = (int)(item - row * nseg) * BXH_SEG;
    const uint8_t* line = im.dst + row * w;
    const int gx = x0 - halo + t;
    if (t < L) buf[0][t] = line[min(max(gx, 0), w - 1)];
    __syncthreads();
    int cur = 0;
    for (int p = 1; p <= 3; ++p) {
      const int lo = (r + 1) * p, hi = L - (r + 1) * p;
      if (t >= lo && t < hi) buf[cur ^ 1][t] = box_tap(buf[cur], t, 1, r, ww, fw);
      __syncthreads();
      if (t < L && (gx < 0 || gx >= w)) buf[cur ^ 1][t] = buf[cur ^ 1][min(max(gx, 0), w - 1) - (x0 - halo)];
      __syncthreads();
      cur ^= 1;
    }
    if (t >= halo && t < halo + BXH_SEG && gx < w) im.tmp[row * w + gx] = buf[cur][t];
    __syncthreads();
  }
}

// vertical: tmp -> dst
__global__ void __launch_bounds__(256) aug_box3_v_kernel(const AugImage* __restrict__ tab) {
  const AugImage& im = tab[blockIdx.y];
  if (im.blur_r < 0 || im.blur_r > BOX_FUSED_R) return;
  __shared__ uint8_t buf[2][(BXV_TH + 6 * (BOX_FUSED_R + 1)) * BXV_TW];
  const int r = im.blur_r, halo = 3 * (r + 1), LH = BXV_TH + 2 * halo, h = im.h, w = im.w;
  const unsigned long long ww = im.blur_ww, fw = im.blur_fw;
  const int ty = (h + BXV_TH - 1) / BXV_TH, tx = (w + BXV_TW - 1) / BXV_TW;
  const long long items = 3ll * ty * tx;
  const int col = threadIdx.x & (BXV_TW - 1), rg = threadIdx.x >> 6;      // 4 row groups
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int c = (int)(item / ((long long)ty * tx));
    const int rem = (int)(item - (long long)c * ty * tx);
    const int y0 = (rem / tx) * BXV_TH, x = (rem % tx) * BXV_TW + col;
    const bool in_x = x < w;
    const uint8_t* plane = im.tmp + (size_t)c * h * w;
    for (int j = rg; j < LH; j += 4)
      buf[0][j * BXV_TW + col] = in_x ? plane[(size_t)min(max(y0 - halo + j, 0), h - 1) * w + x] : 0;
    __syncthreads();
    int cur = 0;
    for (int p = 1; p <= 3; ++p) {
      const int lo = (r + 1) * p, hi = LH - (r + 1) * p;
      for (int j = lo + rg; j < hi; j += 4) buf[cur ^ 1][j * BXV_TW + col] = box_tap(buf[cur] + col, j, BXV_TW, r, ww, fw);
      __syncthreads();
      for (int j = rg; j < LH; j += 4) {
        const int gy = y0 - halo + j;
        if (gy < 0 || gy >= h) buf[cur ^ 1][j * BXV_TW + col] = buf[cur ^ 1][(min(max(gy, 0), h - 1) - (y0 - halo)) * BXV_TW + col];
      }
      __syncthreads();
      cur ^= 1;
    }
    for (int j = halo + rg; j < halo + BXV_TH; j += 4) {
      const int gy = y0 - halo + j;
      if (in_x && gy < h) im.dst[(size_t)c * h * w + (size_t)gy * w + x] = buf[cur][j * BXV_TW + col];
    }
    __syncthreads();
  }
}

__device__ __forceinline__ unsigned int mix32(unsigned int x) {
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}

// RandomErasing(value="random") regions, later ones on top; value = Tensor.byte() of 255 * N(0, 1): truncate, low 8 bits
__global__ void __launch_bounds__(256) aug_erase_kernel(const AugImage* __restrict__ tab, int k) {
  const AugImage& im = tab[blockIdx.y];
  if (k >= im.n_erase) return;
  {
    const int eh = im.eh[k], ew = im.ew[k];
    const size_t n = (size_t)3 * eh * ew, hw = (size_t)im.h * im.w;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      const int c = (int)(i / ((size_t)eh * ew));
      const int rem = (int)(i - (size_t)c * eh * ew);
      const int yy = rem / ew, xx = rem - yy * ew;
      float v;
      if (im.noise[k]) {
        v = im.noise[k][i];
      } else {
        const unsigned int a = mix32(im.seed ^ (unsigned int)(k * 0x9E3779B9u) ^ mix32((unsigned int)i * 2u + 1u));
        const unsigned int b = mix32(a ^ 0x85EBCA6Bu);
        const float u1 = ((float)(a >> 8) + 1.f) * (1.f / 16777216.f), u2 = (float)(b >> 8) * (1.f / 16777216.f);
        v = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
      }
      const float s = __fmul_rn(v, 255.f);
      im.dst[c * hw + (size_t)(im.ei[k] + yy) * im.w + im.ej[k] + xx] = (uint8_t)((long long)truncf(s) & 0xFF);
    }
  }
}

// ------------------------------------------------------------------------------------ weak augmentation: resize + flip
// Pillow Image.resize(..., BILINEAR) (libImaging/Resample.c): separable antialiased convolution, coefficients computed in
// double, normalised, rounded to 22-bit fixed point; uint8 between the horizontal and the vertical pass. The double
// arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction), so the integer coefficients equal Pillow's.
constexpr int RS_BITS = 32 - 8 - 2;

// table layout per axis: bounds[out][2] then kk[out][ksize]
__global__ void resample_coeffs_kernel(int in_size, int out_size, int ksize, int* __restrict__ bounds, int* __restrict__ kk) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx >= out_size) return;
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;                     // bilinear: support 1.0 * filterscale
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dmul_rn((double)xx + 0.5, scale);
  int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) {
    const double t = fabs(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
    ww = __dadd_rn(ww, t < 1.0 ? __dsub_rn(1.0, t) : 0.0);
  }
  int* k = kk + (size_t)xx * ksize;
  for (int x = 0; x < ksize; ++x) {
    int v = 0;
    if (x < xmax) {
      const double t = fabs(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
      double w = t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
      if (ww != 0.0) w = __ddiv_rn(w, ww);
      v = w < 0.0 ? (int)__dadd_rn(-0.5, __dmul_rn(w, (double)(1 << RS_BITS))) : (int)__dadd_rn(0.5, __dmul_rn(w, (double)(1 << RS_BITS)));
    }
    k[x] = v;
  }
  bounds[2 * xx] = xmin;
  bounds[2 * xx + 1] = xmax;
}

// horizontal pass: src uint8 [h, w, 3] (HWC, what image decoders produce) -> tmp uint8 [h, new_w, 3]
__global__ void __launch_bounds__(256) resample_h_kernel(const uint8_t* __restrict__ src, int h, int w, int new_w, int ksize,
                                                         const int* __restrict__ bounds, const int* __restrict__ kk,
                                                         uint8_t* __restrict__ tmp) {
  const size_t n = (size_t)h * new_w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / new_w), xx = (int)(i - (size_t)y * new_w);
    const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
    const int* k = kk + (size_t)xx * ksize;
    const uint8_t* line = src + ((size_t)y * w + xmin) * 3;
    int s0 = 1 << (RS_BITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < xmax; ++x) {
      const int kv = k[x];
      s0 += line[3 * x] * kv; s1 += line[3 * x + 1] * kv; s2 += line[3 * x + 2] * kv;
    }
    uint8_t* o = tmp + i * 3;
    o[0] = (uint8_t)min(max(s0 >> RS_BITS, 0), 255); o[1] = (uint8_t)min(max(s1 >> RS_BITS, 0), 255); o[2] = (uint8_t)min(max(s2 >> RS_BITS, 0), 255);
  }
}

// vertical pass + HWC -> CHW (+ horizontal flip): tmp uint8 [h, new_w, 3] -> dst uint8 [3, new_h, new_w]
__global__ void __launch_bounds__(256) resample_v_kernel(const uint8_t* __restrict__ tmp, int h, int new_w, int new_h, int ksize,
                                                         const int* __restrict__ bounds, const int* __restrict__ kk, int flip,
                                                         uint8_t* __restrict__ dst) {
  const size_t n = (size_t)new_h * new_w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int yy = (int)(i / new_w), xx = (int)(i - (size_t)yy * new_w);
    const int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
    const int* k = kk + (size_t)yy * ksize;
    const uint8_t* col = tmp + ((size_t)ymin * new_w + xx) * 3;
    int s0 = 1 << (RS_BITS - 1), s1 = s0, s2 = s0;
    for (int y = 0; y < ymax; ++y) {
      const int kv = k[y];
      const uint8_t* px = col + (size_t)y * new_w * 3;
      s0 += px[0] * kv; s1 += px[1] * kv; s2 += px[2] * kv;
    }
    const size_t o = (size_t)yy * new_w + (flip ? new_w - 1 - xx : xx);
    dst[o] = (uint8_t)min(max(s0 >> RS_BITS, 0), 255);
    dst[n + o] = (uint8_t)min(max(s1 >> RS_BITS, 0), 255);
    dst[2 * n + o] = (uint8_t)min(max(s2 >> RS_BITS, 0), 255);
  }
}

static int resample_ksize(int in_size, int out_size) {
  const double scale = (double)in_size / out_size;
  const double fs = scale < 1.0 ? 1.0 : scale;
  return (int)ceil(fs) * 2 + 1;
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

// Weak augmentation of the two-crop pipeline (dataset_mapper.py:88-91 -> [D2] ResizeShortestEdge -> ResizeTransform ->
// PIL Image.resize(BILINEAR), [D2] RandomFlip -> HFlipTransform): src uint8 [h, w, 3] HWC -> dst uint8 [3, new_h, new_w]
// CHW, bit-exact with Pillow. tmp: uint8 [h, new_w, 3]; ws: int32 scratch of ut2_resize_workspace_bytes.
extern "C" long long ut2_resize_workspace_bytes(int h, int w, int new_h, int new_w) {
  if (h <= 0 || w <= 0 || new_h <= 0 || new_w <= 0) return 0;
  return 4ll * ((long long)new_w * (2 + resample_ksize(w, new_w)) + (long long)new_h * (2 + resample_ksize(h, new_h))) + 256;
}

extern "C" int ut2_resize_flip_u8(const void* src_hwc, int h, int w, void* dst_chw, int new_h, int new_w, int flip, void* tmp_hwc,
                                  void* ws, long long ws_bytes, void* stream) {
  if (!src_hwc || !dst_chw || !tmp_hwc || !ws) return ut2_fail(-1, "resize: null pointer");
  if (h <= 0 || w <= 0 || new_h <= 0 || new_w <= 0) return ut2_fail(-2, "resize: bad size");
  if (ws_bytes < ut2_resize_workspace_bytes(h, w, new_h, new_w)) return ut2_fail(-3, "resize: workspace too small");
  const int kw = resample_ksize(w, new_w), kh = resample_ksize(h, new_h);
  int* bw = static_cast<int*>(ws);
  int* kkw = bw + 2 * new_w;
  int* bh = kkw + (size_t)new_w * kw;
  int* kkh = bh + 2 * new_h;
  resample_coeffs_kernel<<<(new_w + 127) / 128, 128, 0, STREAM>>>(w, new_w, kw, bw, kkw);
  resample_coeffs_kernel<<<(new_h + 127) / 128, 128, 0, STREAM>>>(h, new_h, kh, bh, kkh);
  const size_t n1 = (size_t)h * new_w, n2 = (size_t)new_h * new_w;
  const int g1 = (int)((n1 + 255) / 256 < 148 * 8 ? (n1 + 255) / 256 : 148 * 8), g2 = (int)((n2 + 255) / 256 < 148 * 8 ? (n2 + 255) / 256 : 148 * 8);
  resample_h_kernel<<<g1, 256, 0, STREAM>>>(static_cast<const uint8_t*>(src_hwc), h, w, new_w, kw, bw, kkw, static_cast<uint8_t*>(tmp_hwc));
  resample_v_kernel<<<g2, 256, 0, STREAM>>>(static_cast<const uint8_t*>(tmp_hwc), h, new_w, new_h, kh, bh, kkh, flip,
                                           static_cast<uint8_t*>(dst_chw));
  return ut2_check_launch("resize_flip");
}

// table: DEVICE array of N 168-byte AugImage records (layout above / in include/ut2.h); lsum_ws: device uint64 [N, 4]
// scratch. Overlapping erase regions are applied by separate launches so that "later on top" holds across blocks.
extern "C" int ut2_strong_augment_u8(const void* table, int N, int max_pixels, int max_erase, unsigned long long* lsum_ws,
                                     void* stream) {
  const int general_blur = max_erase >> 8;      // bit 8 of max_erase: some image has a box radius > 2 (host knows the draws)
  max_erase &= 255;
  if (N <= 0) return 0;
  if (!table || !lsum_ws) return ut2_fail(-1, "strong_augment: null pointer");
  if (max_erase < 0 || max_erase > 3) return ut2_fail(-2, "strong_augment: at most 3 erase regions");
  const AugImage* tab = static_cast<const AugImage*>(table);
  int bx = (max_pixels + 256 * 8 - 1) / (256 * 8);
  if (bx < 1) bx = 1;
  if (bx > 592) bx = 592;
  const dim3 grid(bx, N), grid3(bx * 3 > 592 ? 592 : bx * 3, N);
  cudaError_t e = cudaMemsetAsync(lsum_ws, 0, sizeof(unsigned long long) * 4 * N, STREAM);
  if (e != cudaSuccess) return ut2_fail((int)e, "strong_augment: memset failed");
  aug_copy_kernel<<<grid3, 256, 0, STREAM>>>(tab);
  for (int slot = 0; slot < 4; ++slot) {
    aug_lsum_kernel<<<grid, 256, 0, STREAM>>>(tab, slot, lsum_ws);
    aug_color_kernel<<<grid, 256, 0, STREAM>>>(tab, slot, lsum_ws);
  }
  aug_gray_kernel<<<grid, 256, 0, STREAM>>>(tab);
  aug_box3_h_kernel<<<grid3, 256, 0, STREAM>>>(tab);              // box radius <= 2: three passes per direction in one kernel
  aug_box3_v_kernel<<<grid3, 256, 0, STREAM>>>(tab);
  if (general_blur)                                             // larger radii (never drawn by the reference's [0.1, 2.0])
    for (int pass = 0; pass < 6; ++pass) aug_box_kernel<<<grid3, 256, 0, STREAM>>>(tab, pass, pass >= 3);
  for (int k = 0; k < max_erase; ++k) aug_erase_kernel<<<grid, 256, 0, STREAM>>>(tab, k);   // one launch per region: later ones on top
  return ut2_check_launch("strong_augment");
}
