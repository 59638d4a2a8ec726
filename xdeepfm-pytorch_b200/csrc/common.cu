// error state + device queries for libxdfm_sm100a.so
#include "common.cuh"
#include <stdarg.h>
#include "../../include/xdfm.h"

static thread_local char g_err[1024] = "";
long long g_xdfm_launches = 0;

void xdfm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* xdfm_last_error(void) { return g_err; }
extern "C" int xdfm_version(void) { return 100; }
extern "C" int xdfm_device_cc(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
  return major * 10 + minor;
}
// number of kernels of this library launched so far in this process (bench.py reports the per-step count)
extern "C" long long xdfm_launch_count(void) { return g_xdfm_launches; }
