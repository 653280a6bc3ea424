// C-ABI plumbing shared by all kernels: thread-local last-error string and version/arch queries.
#include <stdarg.h>
#include <stdio.h>
#include "common.cuh"

static thread_local char g_last_error[1024] = "";

void mvlt_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

extern "C" const char* mvlt_last_error(void) { return g_last_error; }

extern "C" int mvlt_abi_version(void) { return 1; }

// 0 if the current device can run the sm_100a kernels, negative otherwise (no CPU fallback exists).
extern "C" int mvlt_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    mvlt_set_error("no CUDA device: %s", cudaGetErrorString(e));
    return (int)e;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    mvlt_set_error("mvlt_b200 kernels are built for sm_100a only; device is sm_%d%d", major, minor);
    return MVLT_ERR_ARG;
  }
  return 0;
}
