// libgode: error state, device info.
#include "common.cuh"
#include <string.h>
#include <vector>
#include <utility>

namespace gode {
static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

struct ProfPair { cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfPair> g_prof[GODE_PROF_KINDS];
static size_t g_prof_used[GODE_PROF_KINDS] = {0};

ProfScope::ProfScope(int kind_, cudaStream_t st_) : kind(kind_), st(st_), slot(nullptr) {
  if (!g_prof_on || kind < 0 || kind >= GODE_PROF_KINDS) return;
  std::vector<ProfPair>& v = g_prof[kind];
  if (g_prof_used[kind] == v.size()) {
    if (v.size() >= 65536) return;
    ProfPair p;
    if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return;
    v.push_back(p);
  }
  ProfPair* p = &v[g_prof_used[kind]++];
  slot = p;
  cudaEventRecord(p->a, st);
}

ProfScope::~ProfScope() {
  if (slot) cudaEventRecord(static_cast<ProfPair*>(slot)->b, st);
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
static int g_reserved_sms = 0;
int persistent_ctas() {
  const int n = sm_count() - g_reserved_sms;
  return n < 1 ? 1 : n;
}
}  // namespace gode

extern "C" int gode_reserve_sms(int n) {
  if (n < 0 || n >= gode::sm_count()) {
    gode::set_error("reserve_sms: %d is outside [0, %d)", n, gode::sm_count());
    return GODE_EINVAL;
  }
  gode::g_reserved_sms = n;
  return GODE_OK;
}

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

extern "C" int gode_profile_enable(int on) {
  gode::g_prof_on = on != 0;
  for (int k = 0; k < GODE_PROF_KINDS; ++k) gode::g_prof_used[k] = 0;
  return GODE_OK;
}

extern "C" int gode_profile_read(int kind, int* launches, float* total_ms, float* max_ms) {
  GODE_REQUIRE(kind >= 0 && kind < GODE_PROF_KINDS, "profile_read: bad kind");
  float tot = 0.f, mx = 0.f;
  int n = 0;
  for (size_t i = 0; i < gode::g_prof_used[kind]; ++i) {
    gode::ProfPair& p = gode::g_prof[kind][i];
    GODE_CHECK_CUDA(cudaEventSynchronize(p.b));
    float ms = 0.f;
    GODE_CHECK_CUDA(cudaEventElapsedTime(&ms, p.a, p.b));
    tot += ms;
    if (ms > mx) mx = ms;
    ++n;
  }
  if (launches) *launches = n;
  if (total_ms) *total_ms = tot;
  if (max_ms) *max_ms = mx;
  return GODE_OK;
}
