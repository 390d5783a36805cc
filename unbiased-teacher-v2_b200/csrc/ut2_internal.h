// Internal helpers shared by the C-ABI translation units (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>

// Records a message retrievable through ut2_last_error_string() and returns `code`.
int ut2_fail(int code, const char* msg);
// cudaGetLastError() after a launch: 0 on success, the cudaError_t (> 0) otherwise.
int ut2_check_launch(const char* what);

static inline int ut2_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// SMs the persistent kernels may fill: the device's count, or the limit set through ut2_set_sm_limit() (ut2_core.cu).
int ut2_sm_budget(int device_sms);

// 3x3 / stride 1 / pad 1 convolutions with 64, 80 or 128 output channels on 2-D patches (conv3x3_halo.cu): 1 = launched,
// 0 = not eligible (the caller falls back to the im2col kernel), < 0 = error.
namespace ut2 {
int conv3x3_halo_try(const void* x, int num_levels, const int* hw, int N, int Cin, const void* w, int Cout, const float* shift,
                     const void* relu_mask, int relu, void* y, int sm_budget, void* stream);
}

// Programmatic dependent launch (PDL): the kernel may become resident while its predecessor on the stream drains — its
// prologue (barrier init, TMEM allocation, descriptor prefetch) overlaps the predecessor's tail — and must execute
// griddep_wait() (sm100_ptx.cuh) before its first global-memory access. Captured into a CUDA graph the attribute becomes a
// programmatic kernel->kernel edge. The 2-3 us it hides per launch matter for short kernels (2+2 images per GPU: -0.3 to -0.4 ms
// of a 13.7 ms step); at 8+8 the board is power-capped and the attribute on every launch measured +0.3 ms (lower clocks), so a
// launch asks for it with an estimate of its duration: on below UT2_PDL_US (default 80 us: 8+8 neutral within noise, 2+2 keeps the
// gain; tools/bench_pdl.sh); UT2_PDL=0 turns it off everywhere.
int ut2_pdl_enabled(double est_us);
// Duration estimate of a launch from its arithmetic and its minimum traffic (1.2 PFLOP/s, 4 TB/s).
static inline double ut2_est_us(double flops, double bytes) {
  const double tf = flops / 1.2e9, tb = bytes / 4.0e6;
  return tf > tb ? tf : tb;
}

#ifdef __CUDACC__
#include <utility>
template <typename... KArgs, typename... Args>
static inline cudaError_t ut2_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                         double est_us, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ut2_pdl_enabled(est_us) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif
