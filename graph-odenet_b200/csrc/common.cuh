// Shared helpers for libgode (sm_100a).  Internal header.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/gode.h"

namespace gode {

void set_error(const char* fmt, ...);

#define GODE_CHECK_CUDA(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::gode::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));  \
      return (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver) ? GODE_ENODEV       \
                                                                             : GODE_ECUDA;      \
    }                                                                                           \
  } while (0)

#define GODE_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::gode::set_error(__VA_ARGS__);      \
      return GODE_EINVAL;                  \
    }                                      \
  } while (0)

extern unsigned long long g_launches;  // kernels launched by this library (for bench.py's gpu_launches)
#define GODE_LAUNCH_CHECK()   \
  do {                        \
    ++::gode::g_launches;     \
    GODE_CHECK_CUDA(cudaGetLastError()); \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int sm_count();
int persistent_ctas();   // sm_count() minus the SMs left free for a concurrent collective (gode_reserve_sms)

// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline leg).
struct ProfScope {
  int kind;
  cudaStream_t st;
  void* slot;
  ProfScope(int kind, cudaStream_t st);
  ~ProfScope();
};

// bump allocator over a caller-provided workspace
struct Arena {
  char* base;
  size_t cap, off;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), cap(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    if (off + bytes > cap) return nullptr;
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
};

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ float4 ld_ro4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace gode
