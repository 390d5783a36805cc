// Library-wide C-ABI plumbing: version, error string, launch checking.
#include <cstdlib>
#include "ut2_internal.h"
#include <stdio.h>
#include <string.h>

static thread_local char g_err[512] = "";

int ut2_fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int ut2_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}

extern "C" const char* ut2_last_error_string(void) { return g_err; }
extern "C" int ut2_version(void) { return 100; }

extern "C" int ut2_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}

// SM budget of the persistent / one-wave kernels (conv forward + data-gradient, weight-gradient split-K). The trainer lowers
// it while gradient all-reduces run on a side stream: those kernels size their grids to the device, and a grid that fills
// all SMs cannot finish before the collective's CTAs release theirs. 0 = the whole device.
static int g_sm_limit = 0;
extern "C" int ut2_set_sm_limit(int n) {
  g_sm_limit = n > 0 ? n : 0;
  return 0;
}
int ut2_pdl_enabled(double est_us) {
  static const double limit_us = [] {
    const char* e = getenv("UT2_PDL");
    if (e && e[0] == '0') return -1.0;
    const char* u = getenv("UT2_PDL_US");
    return u ? atof(u) : 80.0;
  }();
  return est_us < limit_us ? 1 : 0;
}
int ut2_sm_budget(int device_sms) { return (g_sm_limit > 0 && g_sm_limit < device_sms) ? g_sm_limit : device_sms; }
