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
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
    unsigned int r = c0[i], g = c1[i], b = c2[i];
    if (op == 3) {
      hue_px(r, g, b, (unsigned int)im.hue_shift);
    } else {
      const int d = op == 0 ? 0 : (op == 1 ? mean : (int)luma(r, g, b));
      r = blend1(d, (int)r, f, interp); g = blend1(d, (int)g, f, interp); b = blend1(d, (int)b, f, interp);
    }
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

// One extended-box-blur pass (libImaging/BoxBlur.c): out = (ww * sum_{|k|<=r} in[x+k] + fw * (in[x-r-1] + in[x+r+1]) +
// 2^23) >> 24 with replicated edges, along x (vertical = 0) or y (vertical = 1). pass parity picks the ping-pong side.
__global__ void __launch_bounds__(256) aug_box_kernel(const AugImage* __restrict__ tab, int pass, int vertical) {
  const AugImage& im = tab[blockIdx.y];
  if (im.blur_r < 0) return;
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

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

// table: DEVICE array of N 168-byte AugImage records (layout above / in include/ut2.h); lsum_ws: device uint64 [N, 4]
// scratch. Overlapping erase regions are applied by separate launches so that "later on top" holds across blocks.
extern "C" int ut2_strong_augment_u8(const void* table, int N, int max_pixels, int max_erase, unsigned long long* lsum_ws,
                                     void* stream) {
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
  for (int pass = 0; pass < 6; ++pass) aug_box_kernel<<<grid3, 256, 0, STREAM>>>(tab, pass, pass >= 3);
  for (int k = 0; k < max_erase; ++k) aug_erase_kernel<<<grid, 256, 0, STREAM>>>(tab, k);   // one launch per region: later ones on top
  return ut2_check_launch("strong_augment");
}
