// libgode: error state, device info.
#include "common.cuh"
#include <string.h>

namespace gode {
static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}
}  // namespace gode

extern "C" int gode_version(void) { return 100; }

extern "C" unsigned long long gode_launch_count(void) { return gode::g_launches; }

extern "C" const char* gode_last_error(void) { return gode::g_err; }

extern "C" int gode_device_info(int* sm, int* major, int* minor) {
  int dev = 0;
  GODE_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  GODE_CHECK_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm) *sm = p.multiProcessorCount;
  if (major) *major = p.major;
  if (minor) *minor = p.minor;
  return GODE_OK;
}
