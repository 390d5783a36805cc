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
