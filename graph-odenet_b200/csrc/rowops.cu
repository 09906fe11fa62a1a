// libgode: row-local and reduction kernels around the ODE function -- GroupNorm forward/backward,
// column sums, Runge-Kutta linear combinations and the adaptive-step error norm.  All HBM-bound streams.
//
// Reference call sites: nn.GroupNorm(min(32,d), d) GCN/models.py:165,175 (and its autograd);
// bias gradient of GCN/layers.py:35; torchdiffeq rk_common/_compute_error_ratio (restated in oracle/odeint.py).
#include "internal.cuh"
#include <string.h>

namespace gode {

constexpr int kMaxBlocks = 1024;  // partial rows in a column-reduction workspace

// ------------------------------------------------------------------------------------------------
// generic fixed-order reduction of per-block partials:  out[c] = sum_b part[b][c]
// ------------------------------------------------------------------------------------------------
__global__ void k_reduce_partials(int nblk, int width, const float* __restrict__ part, float* __restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= width) return;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += part[(size_t)b * width + c];
  out[c] = s;
}

static int reduce_partials(int nblk, int width, const float* part, float* out, cudaStream_t st) {
  k_reduce_partials<<<(width + 127) / 128, 128, 0, st>>>(nblk, width, part, out);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

static int blocks_for_rows(int64_t n, int rows_per_pass) {
  int64_t want = (n + rows_per_pass - 1) / rows_per_pass;
  int64_t cap = 4LL * sm_count();
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}

// ------------------------------------------------------------------------------------------------
// column sum
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_colsum(int64_t n, int d, const float* __restrict__ x, int64_t ldx,
                                                float* __restrict__ part) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t rows_per_blk = (n + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = blockIdx.x * rows_per_blk, r1 = min(n, r0 + rows_per_blk);
  for (int cb = 0; cb < d; cb += 32) {
    const int c = cb + tx;
    float s = 0.f;
    if (c < d)
      for (int64_t r = r0 + ty; r < r1; r += 8) s += __ldcs(x + r * ldx + c);
    sm[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < d) {
      float t = sm[0][tx];
#pragma unroll
      for (int i = 1; i < 8; ++i) t += sm[i][tx];
      part[(size_t)blockIdx.x * d + c] = t;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm: one thread per (row, group); CPG = channels per group
// ------------------------------------------------------------------------------------------------
template <int CPG>
__device__ __forceinline__ void load_grp(const float* __restrict__ p, float (&v)[CPG], bool vec) {
  if (CPG % 4 == 0 && vec) {
#pragma unroll
    for (int i = 0; i < CPG; i += 4) {
      float4 t = __ldcs(reinterpret_cast<const float4*>(p + i));
      v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < CPG; ++i) v[i] = __ldcs(p + i);
  }
}

template <int CPG>
__device__ __forceinline__ void store_grp(float* __restrict__ p, const float (&v)[CPG], bool vec) {
  if (CPG % 4 == 0 && vec) {
#pragma unroll
    for (int i = 0; i < CPG; i += 4) __stcs(reinterpret_cast<float4*>(p + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
  } else {
#pragma unroll
    for (int i = 0; i < CPG; ++i) __stcs(p + i, v[i]);
  }
}

template <int CPG>
__device__ __forceinline__ void grp_stats(const float (&v)[CPG], float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CPG; ++i) s += v[i];
  mean = s * (1.0f / CPG);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < CPG; ++i) {
    const float dlt = v[i] - mean;
    q += dlt * dlt;
  }
  rstd = 1.0f / sqrtf(q * (1.0f / CPG) + eps);
}

template <int CPG>
__global__ void __launch_bounds__(256) k_gn_fwd(int64_t n, int groups, float eps, const float* __restrict__ x, int64_t ldx,
                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                float* __restrict__ y, int64_t ldy, bool vec) {
  const int64_t total = n * groups;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / groups;
    const int g = static_cast<int>(i - r * groups);
    float v[CPG], o[CPG];
    load_grp<CPG>(x + r * ldx + g * CPG, v, vec);
    float mean, rstd;
    grp_stats<CPG>(v, eps, mean, rstd);
#pragma unroll
    for (int c = 0; c < CPG; ++c) {
      // ATen: y = x * (gamma*rstd) + (beta - mean*gamma*rstd)
      const float sc = __ldg(gamma + g * CPG + c) * rstd;
      o[c] = v[c] * sc + (__ldg(beta + g * CPG + c) - mean * sc);
    }
    store_grp<CPG>(y + r * ldy + g * CPG, o, vec);
  }
}

struct KList {
  const float* k[GODE_MAX_STAGES];
  float c[GODE_MAX_STAGES];
  int n;
};

// Runge-Kutta combination fused behind the GroupNorm backward (the adjoint state of the next augmented stage):
//   anext = a0 + sum_j c[j] k[j] + c_self * dx       (dx itself is stored only when a later stage reads it)
struct RkTail {
  const float* a0;
  KList kl;
  float c_self;
  float* anext;
  gode_rk_second_t second;   // out2 = a0 + sum_j second.coef[j] k[j] + second.coef_self * dx   (out NULL: off)
};

// backward: dx and per-block partial sums of dgamma / dbeta.  blockDim.x is a multiple of `groups`, so a
// thread keeps the same group (and its CPG channels) for every row it visits.
template <int CPG>
__global__ void __launch_bounds__(256) k_gn_bwd(int64_t n, int groups, float eps, const float* __restrict__ x, int64_t ldx,
                                                const float* __restrict__ gamma, const float* __restrict__ dy,
                                                int64_t lddy, float* __restrict__ dx, int64_t lddx,
                                                float* __restrict__ part /*[gridDim][2*d]*/, bool vec, RkTail rk) {
  extern __shared__ float sm[];  // [blockDim][2*CPG]
  const int64_t total = n * groups;
  const int g = threadIdx.x % groups;
  float gam[CPG], dg[CPG], db[CPG];
#pragma unroll
  for (int c = 0; c < CPG; ++c) {
    gam[c] = __ldg(gamma + g * CPG + c);
    dg[c] = 0.f;
    db[c] = 0.f;
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / groups;
    float v[CPG], u[CPG], o[CPG];
    load_grp<CPG>(x + r * ldx + g * CPG, v, vec);
    load_grp<CPG>(dy + r * lddy + g * CPG, u, vec);
    float mean, rstd;
    grp_stats<CPG>(v, eps, mean, rstd);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int c = 0; c < CPG; ++c) {
      const float xh = (v[c] - mean) * rstd;
      const float h = u[c] * gam[c];
      m1 += h;
      m2 += h * xh;
      dg[c] += u[c] * xh;
      db[c] += u[c];
      v[c] = xh;
      u[c] = h;
    }
    m1 *= (1.0f / CPG);
    m2 *= (1.0f / CPG);
#pragma unroll
    for (int c = 0; c < CPG; ++c) o[c] = rstd * (u[c] - m1 - v[c] * m2);
    if (dx) store_grp<CPG>(dx + r * lddx + g * CPG, o, vec);
    if (rk.anext || rk.second.out) {   // contiguous [n, d] operands (ld = d)
      const int64_t off = r * (int64_t)(groups * CPG) + g * CPG;
      float acc[CPG], acc2[CPG], w[CPG];
      load_grp<CPG>(rk.a0 + off, acc, vec);
#pragma unroll
      for (int c = 0; c < CPG; ++c) {
        acc2[c] = acc[c] + rk.second.coef_self * o[c];
        acc[c] += rk.c_self * o[c];
      }
#pragma unroll
      for (int j = 0; j < GODE_MAX_STAGES; ++j)
        if (j < rk.kl.n) {
          load_grp<CPG>(rk.kl.k[j] + off, w, vec);
#pragma unroll
          for (int c = 0; c < CPG; ++c) {
            acc[c] += rk.kl.c[j] * w[c];
            acc2[c] += rk.second.coef[j] * w[c];
          }
        }
      if (rk.anext) store_grp<CPG>(rk.anext + off, acc, vec);
      if (rk.second.out) store_grp<CPG>(rk.second.out + off, acc2, vec);
    }
  }
#pragma unroll
  for (int c = 0; c < CPG; ++c) {
    sm[threadIdx.x * 2 * CPG + c] = dg[c];
    sm[threadIdx.x * 2 * CPG + CPG + c] = db[c];
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    const int d = groups * CPG;
#pragma unroll
    for (int c = 0; c < CPG; ++c) {
      float a = 0.f, b = 0.f;
      for (int t = threadIdx.x; t < blockDim.x; t += groups) {
        a += sm[t * 2 * CPG + c];
        b += sm[t * 2 * CPG + CPG + c];
      }
      part[(size_t)blockIdx.x * 2 * d + g * CPG + c] = a;
      part[(size_t)blockIdx.x * 2 * d + d + g * CPG + c] = b;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Runge-Kutta combination and error norm
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rk_combine4(int64_t n4, const float* __restrict__ y0, KList kl, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < GODE_MAX_STAGES; ++j)
      if (j < kl.n) {
        const float4 k = ld_stream4(kl.k[j] + 4 * i);
        t.x += kl.c[j] * k.x; t.y += kl.c[j] * k.y; t.z += kl.c[j] * k.z; t.w += kl.c[j] * k.w;
      }
    if (y0) {
      const float4 y = ld_stream4(y0 + 4 * i);
      t.x += y.x; t.y += y.y; t.z += y.z; t.w += y.w;
    }
    st_stream4(out + 4 * i, t);
  }
}

__global__ void __launch_bounds__(256) k_rk_combine1(int64_t n, const float* __restrict__ y0, KList kl, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float t = 0.f;
    for (int j = 0; j < kl.n; ++j) t += kl.c[j] * kl.k[j][i];
    out[i] = y0 ? y0[i] + t : t;
  }
}

__global__ void __launch_bounds__(256) k_rk_err(int64_t n, const float* __restrict__ y0, const float* __restrict__ y1, KList kl,
                                                float rtol, float atol, float* __restrict__ part) {
  __shared__ float sm[8];
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float e = 0.f;
#pragma unroll
    for (int j = 0; j < GODE_MAX_STAGES; ++j)
      if (j < kl.n) e += kl.c[j] * __ldcs(kl.k[j] + i);
    const float tol = atol + rtol * fmaxf(fabsf(__ldcs(y0 + i)), fabsf(__ldcs(y1 + i)));
    const float q = e / tol;
    acc += q * q;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sm[i];
    part[blockIdx.x] = t;
  }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int colsum(int64_t n, int d, const float* x, int64_t ldx, float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (ws_bytes < gode_colreduce_workspace_bytes(d) || !ws) {
    set_error("colsum: workspace too small");
    return GODE_EWORKSPACE;
  }
  if (n == 0) {
    GODE_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * d, st));
    return GODE_OK;
  }
  int nblk = blocks_for_rows(n, 64);
  k_colsum<<<nblk, 256, 0, st>>>(n, d, x, ldx, static_cast<float*>(ws));
  GODE_LAUNCH_CHECK();
  return reduce_partials(nblk, d, static_cast<float*>(ws), out, st);
}

template <int CPG>
static int gn_fwd_t(int64_t n, int groups, float eps, const float* x, int64_t ldx, const float* gamma, const float* beta,
                    float* y, int64_t ldy, cudaStream_t st) {
  const bool vec = al16(x) && al16(y) && ldx % 4 == 0 && ldy % 4 == 0;
  int64_t total = n * groups;
  int64_t want = (total + 255) / 256;
  int64_t cap = 32LL * sm_count();
  int grid = static_cast<int>(want < cap ? want : cap);
  if (grid < 1) grid = 1;
  k_gn_fwd<CPG><<<grid, 256, 0, st>>>(n, groups, eps, x, ldx, gamma, beta, y, ldy, vec);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

template <int CPG>
static int gn_bwd_t(int64_t n, int groups, float eps, const float* x, int64_t ldx, const float* gamma, const float* dy,
                    int64_t lddy, float* dx, int64_t lddx, float* dgamma, float* dbeta, float* part, cudaStream_t st,
                    const RkTail& rk) {
  bool vec = al16(x) && al16(dy) && al16(dx) && ldx % 4 == 0 && lddy % 4 == 0 && lddx % 4 == 0;
  if (rk.anext || rk.second.out) {
    vec = vec && al16(rk.a0) && al16(rk.anext) && al16(rk.second.out) && (groups * CPG) % 4 == 0;
    for (int j = 0; j < rk.kl.n; ++j) vec = vec && al16(rk.kl.k[j]);
  }
  const int d = groups * CPG;
  const int tpb = (256 / groups) * groups;
  int64_t total = n * groups;
  int64_t want = (total + tpb - 1) / tpb;
  int64_t cap = 4LL * sm_count();
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  int grid = static_cast<int>(want < cap ? want : cap);
  if (grid < 1) grid = 1;
  k_gn_bwd<CPG><<<grid, tpb, sizeof(float) * tpb * 2 * CPG, st>>>(n, groups, eps, x, ldx, gamma, dy, lddy, dx, lddx, part, vec, rk);
  GODE_LAUNCH_CHECK();
  // partial layout per block: [dgamma d | dbeta d]
  k_reduce_partials<<<(2 * d + 127) / 128, 128, 0, st>>>(grid, 2 * d, part, part + (size_t)kMaxBlocks * 2 * d);
  GODE_LAUNCH_CHECK();
  GODE_CHECK_CUDA(cudaMemcpyAsync(dgamma, part + (size_t)kMaxBlocks * 2 * d, sizeof(float) * d, cudaMemcpyDeviceToDevice, st));
  GODE_CHECK_CUDA(cudaMemcpyAsync(dbeta, part + (size_t)kMaxBlocks * 2 * d + d, sizeof(float) * d, cudaMemcpyDeviceToDevice, st));
  return GODE_OK;
}

int groupnorm_fwd(int64_t n, int d, int groups, float eps, const float* x, int64_t ldx, const float* gamma,
                  const float* beta, float* y, int64_t ldy, cudaStream_t st) {
  if (n == 0) return GODE_OK;
  const int cpg = d / groups;
  switch (cpg) {
    case 1: return gn_fwd_t<1>(n, groups, eps, x, ldx, gamma, beta, y, ldy, st);
    case 2: return gn_fwd_t<2>(n, groups, eps, x, ldx, gamma, beta, y, ldy, st);
    case 4: return gn_fwd_t<4>(n, groups, eps, x, ldx, gamma, beta, y, ldy, st);
    case 8: return gn_fwd_t<8>(n, groups, eps, x, ldx, gamma, beta, y, ldy, st);
    case 16: return gn_fwd_t<16>(n, groups, eps, x, ldx, gamma, beta, y, ldy, st);
    default: set_error("groupnorm: %d channels per group unsupported (1,2,4,8,16)", cpg); return GODE_EINVAL;
  }
}

static int fill_klist(KList& kl, const float* const* k, const float* c, int n);

int groupnorm_bwd_rk(int64_t n, int d, int groups, float eps, const float* x, int64_t ldx, const float* gamma,
                     const float* dy, int64_t lddy, float* dx, int64_t lddx, float* dgamma, float* dbeta, void* ws,
                     size_t ws_bytes, cudaStream_t st, const float* a0, const float* const* kprev, const float* coef,
                     int n_prev, float coef_self, float* a_next, const gode_rk_second_t* second) {
  if (ws_bytes < gode_colreduce_workspace_bytes(2 * d) || !ws) {
    set_error("groupnorm_bwd: workspace too small");
    return GODE_EWORKSPACE;
  }
  float* out2 = second ? second->out : nullptr;
  GODE_REQUIRE(dx || a_next || out2, "groupnorm_bwd: nothing to write");
  GODE_REQUIRE((!a_next && !out2) || a0, "groupnorm_bwd: a_next needs a0");
  if (n == 0) {
    GODE_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * d, st));
    GODE_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * d, st));
    return GODE_OK;
  }
  RkTail rk;
  rk.a0 = a0;
  rk.c_self = coef_self;
  rk.anext = a_next;
  if (second) rk.second = *second; else memset(&rk.second, 0, sizeof(rk.second));
  int rc = fill_klist(rk.kl, kprev, coef, (a_next || out2) ? n_prev : 0);
  if (rc) return rc;
  float* part = static_cast<float*>(ws);
  const int cpg = d / groups;
  switch (cpg) {
    case 1: return gn_bwd_t<1>(n, groups, eps, x, ldx, gamma, dy, lddy, dx, lddx, dgamma, dbeta, part, st, rk);
    case 2: return gn_bwd_t<2>(n, groups, eps, x, ldx, gamma, dy, lddy, dx, lddx, dgamma, dbeta, part, st, rk);
    case 4: return gn_bwd_t<4>(n, groups, eps, x, ldx, gamma, dy, lddy, dx, lddx, dgamma, dbeta, part, st, rk);
    case 8: return gn_bwd_t<8>(n, groups, eps, x, ldx, gamma, dy, lddy, dx, lddx, dgamma, dbeta, part, st, rk);
    case 16: return gn_bwd_t<16>(n, groups, eps, x, ldx, gamma, dy, lddy, dx, lddx, dgamma, dbeta, part, st, rk);
    default: set_error("groupnorm: %d channels per group unsupported (1,2,4,8,16)", cpg); return GODE_EINVAL;
  }
}

int groupnorm_bwd(int64_t n, int d, int groups, float eps, const float* x, int64_t ldx, const float* gamma,
                  const float* dy, int64_t lddy, float* dx, int64_t lddx, float* dgamma, float* dbeta, void* ws,
                  size_t ws_bytes, cudaStream_t st) {
  return groupnorm_bwd_rk(n, d, groups, eps, x, ldx, gamma, dy, lddy, dx, lddx, dgamma, dbeta, ws, ws_bytes, st, nullptr,
                          nullptr, nullptr, 0, 0.f, nullptr, nullptr);
}

static int fill_klist(KList& kl, const float* const* k, const float* c, int n) {
  GODE_REQUIRE(n >= 0 && n <= GODE_MAX_STAGES, "rk: n_k out of range");
  kl.n = n;
  for (int j = 0; j < GODE_MAX_STAGES; ++j) {
    kl.k[j] = j < n ? k[j] : nullptr;
    kl.c[j] = j < n ? c[j] : 0.f;
  }
  return GODE_OK;
}

int rk_combine(int64_t n, const float* y0, const float* const* k, const float* c, int nk, float* out, cudaStream_t st) {
  if (n == 0) return GODE_OK;
  KList kl;
  int rc = fill_klist(kl, k, c, nk);
  if (rc) return rc;
  bool vec = (n % 4 == 0) && al16(y0) && al16(out);
  for (int j = 0; j < nk; ++j) vec = vec && al16(k[j]);
  int64_t work = vec ? n / 4 : n;
  int64_t want = (work + 255) / 256, cap = 16LL * sm_count();
  int grid = static_cast<int>(want < cap ? want : cap);
  if (vec) k_rk_combine4<<<grid, 256, 0, st>>>(n / 4, y0, kl, out);
  else k_rk_combine1<<<grid, 256, 0, st>>>(n, y0, kl, out);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

}  // namespace gode

using namespace gode;

extern "C" size_t gode_colreduce_workspace_bytes(int32_t width) {
  return sizeof(float) * (static_cast<size_t>(kMaxBlocks) + 1) * static_cast<size_t>(width > 0 ? width : 1);
}

extern "C" int gode_colsum_f32(int64_t n, int32_t d, const float* x, int64_t ldx, float* out, void* ws, size_t ws_bytes,
                               void* stream) {
  GODE_REQUIRE(n >= 0 && d > 0 && ldx >= d && out && (n == 0 || x), "colsum: bad argument");
  return colsum(n, d, x, ldx, out, ws, ws_bytes, as_stream(stream));
}

extern "C" int gode_groupnorm_fwd(int64_t n, int32_t d, int32_t groups, float eps, const float* x, int64_t ldx,
                                  const float* gamma, const float* beta, float* y, int64_t ldy, void* stream) {
  GODE_REQUIRE(n >= 0 && d > 0 && groups > 0 && d % groups == 0 && ldx >= d && ldy >= d, "groupnorm_fwd: bad shape");
  GODE_REQUIRE(gamma && beta && (n == 0 || (x && y)), "groupnorm_fwd: null pointer");
  return groupnorm_fwd(n, d, groups, eps, x, ldx, gamma, beta, y, ldy, as_stream(stream));
}

extern "C" int gode_groupnorm_bwd(int64_t n, int32_t d, int32_t groups, float eps, const float* x, int64_t ldx,
                                  const float* gamma, const float* dy, int64_t lddy, float* dx, int64_t lddx,
                                  float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream) {
  GODE_REQUIRE(n >= 0 && d > 0 && groups > 0 && d % groups == 0 && groups <= 256, "groupnorm_bwd: bad shape");
  GODE_REQUIRE(gamma && dgamma && dbeta && (n == 0 || (x && dy && dx)), "groupnorm_bwd: null pointer");
  return groupnorm_bwd(n, d, groups, eps, x, ldx, gamma, dy, lddy, dx, lddx, dgamma, dbeta, ws, ws_bytes, as_stream(stream));
}

extern "C" int gode_rk_combine(int64_t n_elems, const float* y0, const float* const* k_host, const float* coef_host,
                               int32_t n_k, float* out, void* stream) {
  GODE_REQUIRE(n_elems >= 0 && out && (n_k == 0 || (k_host && coef_host)), "rk_combine: bad argument");
  return rk_combine(n_elems, y0, k_host, coef_host, n_k, out, as_stream(stream));
}

extern "C" int gode_rk_error_sumsq(int64_t n_elems, const float* y0, const float* y1, const float* const* k_host,
                                   const float* coef_host, int32_t n_k, float rtol, float atol, float* sumsq_out, void* ws,
                                   size_t ws_bytes, void* stream) {
  GODE_REQUIRE(n_elems >= 0 && y0 && y1 && sumsq_out && k_host && coef_host, "rk_error: bad argument");
  if (ws_bytes < gode_colreduce_workspace_bytes(1) || !ws) {
    set_error("rk_error: workspace too small");
    return GODE_EWORKSPACE;
  }
  KList kl;
  int rc = fill_klist(kl, k_host, coef_host, n_k);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  int64_t want = (n_elems + 1023) / 1024, cap = 4LL * sm_count();
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  int grid = static_cast<int>(want < cap ? want : cap);
  if (grid < 1) grid = 1;
  k_rk_err<<<grid, 256, 0, st>>>(n_elems, y0, y1, kl, rtol, atol, static_cast<float*>(ws));
  GODE_LAUNCH_CHECK();
  return reduce_partials(grid, 1, static_cast<float*>(ws), sumsq_out, st);
}

// res = g * (out > 0): backward of a fused ReLU epilogue (F.relu in GCN/models.py:76-80, MLP hidden layers QC/layers.py:55)
namespace gode {
__global__ void __launch_bounds__(256) k_relu_bwd(int64_t n, const float* __restrict__ g, const float* __restrict__ out,
                                                  float* __restrict__ res) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    res[i] = __ldcs(out + i) > 0.f ? __ldcs(g + i) : 0.f;
}
}  // namespace gode

extern "C" int gode_relu_bwd(int64_t n_elems, const float* g, const float* out, float* res, void* stream) {
  GODE_REQUIRE(n_elems >= 0 && (n_elems == 0 || (g && out && res)), "relu_bwd: bad argument");
  if (n_elems == 0) return GODE_OK;
  int64_t blocks = (n_elems + 255) / 256;
  const int64_t cap = 8LL * sm_count();
  if (blocks > cap) blocks = cap;
  k_relu_bwd<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(n_elems, g, out, res);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}
