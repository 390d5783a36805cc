// Library-wide C-ABI plumbing: version, error string, launch checking.
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
