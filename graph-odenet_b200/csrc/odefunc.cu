// libgode: the GCN ODE function  f(t,y) = relu(A_hat ([t || GroupNorm(y)] W) + b)  and its VJP, composed
// from the gather (spmm.cu), dense (gemm.cu / transform_tc.cu) and row-local (rowops.cu) kernels.
//
// Reference: ODEfunc.forward GCN/models.py:172-179; FixedGraphConvolution.forward GCN/layers.py:69-75;
// the adjoint's torch.autograd.grad over (t, y, theta) restated in oracle/odeint.py:_Adjoint.
//
// The time column is never materialised: [t || z] W = t * W[0,:] + z * W[1:,:].
#include "internal.cuh"
#include <string.h>

namespace gode {

// transform_tc.cu (tcgen05 kernels)
int transform_tc(const gode_gcn_odefunc_t* f, const float* y, float t, float* S, cudaStream_t st, int64_t row0 = 0,
                 int64_t n_rows = -1);
bool transform_tc_supported(const gode_gcn_odefunc_t* f);
int input_grad_tc(const gode_gcn_odefunc_t* f, const float* gS, float* gz, cudaStream_t st);
bool wgrad_tc_supported(const gode_gcn_odefunc_t* f);
bool input_grad_tc_supported(const gode_gcn_odefunc_t* f);
size_t wgrad_tc_ws_bytes(const gode_gcn_odefunc_t* f);
int wgrad_tc(const gode_gcn_odefunc_t* f, const float* y, const float* gS, float* cs, float* gW1, float* ws,
             size_t ws_bytes, cudaStream_t st);

struct GcnWs {
  float* heavy;    // heavy-row partial sums of the SpMMs
  size_t heavy_bytes;
  float* bufA;
  float* bufB;
  float* bufC;
  float* red;      // column-reduction workspace (width 2d)
  size_t red_bytes;
  float* splitk;
  size_t splitk_bytes;
  float* small;    // [d] scratch
};

static int64_t max_rows(const gode_gcn_odefunc_t* f) {
  // scratch tensors are [n_rows, d]; the gather operands (which may be longer: owned + halo rows) are the caller's
  return f->A.n_rows;
}

static size_t heavy_bytes_for(const gode_gcn_odefunc_t* f) {
  size_t a = spmm_ws_bytes(f->A, f->d);
  size_t b = f->At.rowptr ? spmm_ws_bytes(f->At, f->d) : 0;
  return a > b ? a : b;
}

static int splits_for(int64_t k) {
  int64_t s = k / 4096;
  int64_t cap = 2LL * sm_count();
  if (s > cap) s = cap;
  if (s < 1) s = 1;
  return static_cast<int>(s);
}

static size_t splitk_bytes_for(const gode_gcn_odefunc_t* f) {
  size_t a = sizeof(float) * static_cast<size_t>(splits_for(f->A.n_rows)) * f->d * f->d;
  size_t b = wgrad_tc_supported(f) ? wgrad_tc_ws_bytes(f) : 0;
  return a > b ? a : b;
}

static size_t ws_bytes_for(const gode_gcn_odefunc_t* f) {
  const size_t nd = align_up(sizeof(float) * static_cast<size_t>(max_rows(f)) * f->d, 256);
  const size_t red = align_up(gode_colreduce_workspace_bytes(2 * f->d), 256);
  const size_t sk = align_up(splitk_bytes_for(f), 256);
  return 3 * nd + red + sk + heavy_bytes_for(f) + 8192;
}

static int carve(const gode_gcn_odefunc_t* f, void* ws, size_t ws_bytes, GcnWs& w) {
  if (!ws || ws_bytes < ws_bytes_for(f)) {
    set_error("gcn: workspace too small (%zu < %zu)", ws_bytes, ws_bytes_for(f));
    return GODE_EWORKSPACE;
  }
  Arena ar(ws, ws_bytes);
  const size_t nd = static_cast<size_t>(max_rows(f)) * f->d;
  w.heavy_bytes = heavy_bytes_for(f);
  w.heavy = w.heavy_bytes ? reinterpret_cast<float*>(ar.take<char>(w.heavy_bytes)) : nullptr;
  w.bufA = ar.take<float>(nd);
  w.bufB = ar.take<float>(nd);
  w.bufC = ar.take<float>(nd);
  w.red_bytes = gode_colreduce_workspace_bytes(2 * f->d);
  w.red = reinterpret_cast<float*>(ar.take<char>(w.red_bytes));
  w.splitk_bytes = splitk_bytes_for(f);
  w.splitk = reinterpret_cast<float*>(ar.take<char>(w.splitk_bytes));
  w.small = ar.take<float>(1024);
  if (!w.small) {
    set_error("gcn: workspace arena exhausted");
    return GODE_EWORKSPACE;
  }
  return GODE_OK;
}

static int check(const gode_gcn_odefunc_t* f) {
  GODE_REQUIRE(f != nullptr, "gcn: null descriptor");
  GODE_REQUIRE(f->A.n_rows >= 0 && f->d > 0 && f->groups > 0 && f->d % f->groups == 0, "gcn: bad shape");
  GODE_REQUIRE(f->W && f->gamma && f->beta && f->A.rowptr, "gcn: null parameter pointer");
  return GODE_OK;
}

// gW[0,:] = t * cs ; gt = <cs, W[0,:]>      (cs = column sums of gS)
__global__ void k_time_terms(int d, float t, const float* __restrict__ cs, const float* __restrict__ W0,
                             float* __restrict__ gW0, float* __restrict__ gt) {
  __shared__ float sm[32];
  float acc = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float v = cs[c];
    gW0[c] = t * v;
    acc += v * W0[c];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) s += sm[i];
    *gt = s;
  }
}

// The descriptor's per-call second combination (header: gode_gcn_odefunc_t.second) joins the epilogue of this gather; it
// shares y0 and the k_j list with y_next (which may itself be off).
static int set_second(gode_spmm_epilogue_t& ep, const gode_gcn_odefunc_t* f, const float* y0, const float* const* kprev_host,
                      int n_prev) {
  if (!f->second.out) return GODE_OK;
  GODE_REQUIRE(y0, "gcn: the second Runge-Kutta combination needs y0");
  ep.second = f->second;
  ep.y0 = y0;
  ep.n_prev = n_prev;
  for (int j = 0; j < n_prev; ++j) ep.kprev[j] = kprev_host[j];
  return GODE_OK;
}

static int transform_impl(const gode_gcn_odefunc_t* f, const float* y, float t, float* S, GcnWs& w, cudaStream_t st) {
  ProfScope prof(GODE_PROF_TRANSFORM, st);
  if (transform_tc_supported(f)) return transform_tc(f, y, t, S, st);
  GODE_REQUIRE(!f->push_S.ptr, "gcn_transform: a fused halo push needs the tensor-core transform (gode_gcn_push_fusable)");
  const int d = f->d;
  int rc = groupnorm_fwd(f->A.n_rows, d, f->groups, f->gn_eps, y, d, f->gamma, f->beta, w.bufC, d, st);
  if (rc) return rc;
  return gemm_simt(0, 0, f->A.n_rows, d, d, 1.f, w.bufC, d, f->W + d, d, 0.f, S, d, 1, nullptr, 0, st, f->W, t);
}

}  // namespace gode

using namespace gode;

extern "C" size_t gode_gcn_workspace_bytes(const gode_gcn_odefunc_t* f) {
  if (!f || f->d <= 0) return 0;
  return ws_bytes_for(f);
}

extern "C" int gode_gcn_push_fusable(const gode_gcn_odefunc_t* f) {
  if (!f || f->d <= 0 || f->groups <= 0) return 0;
  return transform_tc_supported(f) ? 1 : 0;   // the vectorised gather kernels cover every width the transform does
}

extern "C" int gode_gcn_transform(const gode_gcn_odefunc_t* f, const float* y, float t, float* S, void* ws,
                                  size_t ws_bytes, void* stream) {
  int rc = check(f);
  if (rc) return rc;
  GODE_REQUIRE(f->A.n_rows == 0 || (y && S), "gcn_transform: null pointer");
  GcnWs w;
  rc = carve(f, ws, ws_bytes, w);
  if (rc) return rc;
  return transform_impl(f, y, t, S, w, as_stream(stream));
}

extern "C" int gode_gcn_transform_rows(const gode_gcn_odefunc_t* f, const float* y, float t, float* S, int64_t row0,
                                       int64_t n_rows, void* ws, size_t ws_bytes, void* stream) {
  int rc = check(f);
  if (rc) return rc;
  const int64_t row_limit = (transform_tc_supported(f) && f->A.n_cols > f->A.n_rows) ? f->A.n_cols : f->A.n_rows;
  GODE_REQUIRE(row0 >= 0 && n_rows >= 0 && row0 + n_rows <= row_limit, "gcn_transform_rows: row range outside the block");
  if (n_rows == 0) return GODE_OK;
  GODE_REQUIRE(y && S, "gcn_transform_rows: null pointer");
  GcnWs w;
  rc = carve(f, ws, ws_bytes, w);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  ProfScope prof(GODE_PROF_TRANSFORM, st);
  if (transform_tc_supported(f)) return transform_tc(f, y, t, S, st, row0, n_rows);
  GODE_REQUIRE(!f->push_S.ptr, "gcn_transform_rows: a fused halo push needs the tensor-core transform");
  const int d = f->d;
  const float* yr = y + row0 * d;
  float* Sr = S + row0 * d;
  rc = groupnorm_fwd(n_rows, d, f->groups, f->gn_eps, yr, d, f->gamma, f->beta, w.bufC, d, st);
  if (rc) return rc;
  return gemm_simt(0, 0, n_rows, d, d, 1.f, w.bufC, d, f->W + d, d, 0.f, Sr, d, 1, nullptr, 0, st, f->W, t);
}

extern "C" int gode_gcn_stage_fwd(const gode_gcn_odefunc_t* f, const float* S, float* k_out, const float* y0,
                                  const float* const* kprev_host, const float* coef_host, int32_t n_prev, float coef_self,
                                  float* y_next, float t_next, float* S_next, void* ws, size_t ws_bytes, void* stream) {
  int rc = check(f);
  if (rc) return rc;
  GODE_REQUIRE(n_prev >= 0 && n_prev <= GODE_MAX_STAGES, "gcn_stage_fwd: n_prev out of range");
  GODE_REQUIRE(!S_next || y_next, "gcn_stage_fwd: S_next needs y_next");
  GODE_REQUIRE(!y_next || y0, "gcn_stage_fwd: y_next needs y0");
  cudaStream_t st = as_stream(stream);
  GcnWs w;
  rc = carve(f, ws, ws_bytes, w);
  if (rc) return rc;
  gode_spmm_epilogue_t ep;
  memset(&ep, 0, sizeof(ep));
  ep.bias = f->b;
  ep.relu = 1;
  if (y_next) {
    ep.y0 = y0;
    ep.n_prev = n_prev;
    for (int j = 0; j < n_prev; ++j) {
      ep.kprev[j] = kprev_host[j];
      ep.coef[j] = coef_host[j];
    }
    ep.coef_self = coef_self;
    ep.ynext = y_next;
    ep.push_y = f->push_y;
  }
  if ((rc = set_second(ep, f, y0, kprev_host, n_prev))) return rc;
  {
    ProfScope prof(GODE_PROF_AGG_FWD, st);
    ep.acc_in = f->partial_in;
    rc = spmm_dispatch(f->A, S + f->gather_row_offset * f->d, f->d, f->d, k_out, f->d, ep, w.heavy, w.heavy_bytes, st);
  }
  if (rc) return rc;
  if (S_next) rc = transform_impl(f, y_next, t_next, S_next, w, st);
  return rc;
}

// The same stage for the rows [row0, row0 + n_rows) of the block only (k_out / y_next rows outside the range are not
// touched): the row-partitioned solver gathers a stage in row chunks so that chunk c's rows of y_next can be transformed and
// pushed to the peers while chunk c+1 is still being gathered (parallel.py, GODE_PIPE_G).
extern "C" int gode_gcn_stage_fwd_rows(const gode_gcn_odefunc_t* f, const float* S, float* k_out, const float* y0,
                                       const float* const* kprev_host, const float* coef_host, int32_t n_prev, float coef_self,
                                       float* y_next, int64_t row0, int64_t n_rows, void* ws, size_t ws_bytes, void* stream) {
  int rc = check(f);
  if (rc) return rc;
  GODE_REQUIRE(n_prev >= 0 && n_prev <= GODE_MAX_STAGES, "gcn_stage_fwd_rows: n_prev out of range");
  GODE_REQUIRE(!y_next || y0, "gcn_stage_fwd_rows: y_next needs y0");
  GODE_REQUIRE(row0 >= 0 && n_rows >= 0 && row0 + n_rows <= f->A.n_rows, "gcn_stage_fwd_rows: row range outside the block");
  if (n_rows == 0) return GODE_OK;
  cudaStream_t st = as_stream(stream);
  GcnWs w;
  rc = carve(f, ws, ws_bytes, w);
  if (rc) return rc;
  gode_spmm_epilogue_t ep;
  memset(&ep, 0, sizeof(ep));
  ep.bias = f->b;
  ep.relu = 1;
  if (y_next) {
    ep.y0 = y0;
    ep.n_prev = n_prev;
    for (int j = 0; j < n_prev; ++j) {
      ep.kprev[j] = kprev_host[j];
      ep.coef[j] = coef_host[j];
    }
    ep.coef_self = coef_self;
    ep.ynext = y_next;
    ep.push_y = f->push_y;
  }
  if ((rc = set_second(ep, f, y0, kprev_host, n_prev))) return rc;
  ep.acc_in = f->partial_in;
  ProfScope prof(GODE_PROF_AGG_FWD, st);
  return spmm_dispatch(f->A, S + f->gather_row_offset * f->d, f->d, f->d, k_out, f->d, ep, w.heavy, w.heavy_bytes, st, row0,
                       row0 + n_rows);
}

extern "C" int gode_gcn_vjp_phase1(const gode_gcn_odefunc_t* f, const float* S, const float* a, float sign, float* k_y,
                                   float* gP, const float* y0, const float* const* kprev_host, const float* coef_host,
                                   int32_t n_prev, float coef_self, float* y_next, void* ws, size_t ws_bytes,
                                   void* stream) {
  int rc = check(f);
  if (rc) return rc;
  GODE_REQUIRE(f->A.n_rows == 0 || (S && a && gP), "gcn_vjp_phase1: null pointer");
  GcnWs w;
  rc = carve(f, ws, ws_bytes, w);
  if (rc) return rc;
  GODE_REQUIRE(n_prev >= 0 && n_prev <= GODE_MAX_STAGES && (!y_next || y0), "gcn_vjp_phase1: bad RK arguments");
  gode_spmm_epilogue_t ep;
  memset(&ep, 0, sizeof(ep));
  ep.bias = f->b;
  ep.relu = 1;
  ep.mask_src = a;
  ep.mask_scale = sign;
  ep.gp_out = gP;
  ep.push = f->push_gP;
  if (y_next) {
    ep.y0 = y0;
    ep.ynext = y_next;
    ep.push_y = f->push_y;
    ep.n_prev = n_prev;
    ep.coef_self = coef_self;
    for (int j = 0; j < n_prev; ++j) {
      ep.kprev[j] = kprev_host[j];
      ep.coef[j] = coef_host[j];
    }
  }
  if ((rc = set_second(ep, f, y0, kprev_host, n_prev))) return rc;
  ep.gp_row_scale = f->gp_row_scale;
  ep.acc_in = f->partial_in;
  ProfScope prof(GODE_PROF_AGG_FWD, as_stream(stream));
  return spmm_dispatch(f->A, S + f->gather_row_offset * f->d, f->d, f->d, k_y, f->d, ep, w.heavy, w.heavy_bytes,
                       as_stream(stream));
}

extern "C" int gode_gcn_vjp_phase2_rk(const gode_gcn_odefunc_t* f, const float* y, float t, const float* gP, float* k_a,
                                      float* gtheta, const float* a0, const float* const* kprev_host,
                                      const float* coef_host, int32_t n_prev, float coef_self, float* a_next, void* ws,
                                      size_t ws_bytes, void* stream) {
  int rc = check(f);
  if (rc) return rc;
  GODE_REQUIRE(f->At.rowptr && (f->A.n_rows == 0 || (y && gP && (k_a || a_next || f->second.out))) && gtheta,
               "gcn_vjp_phase2: null pointer");
  GODE_REQUIRE(n_prev >= 0 && n_prev <= GODE_MAX_STAGES && ((!a_next && !f->second.out) || a0),
               "gcn_vjp_phase2: bad RK arguments");
  GODE_REQUIRE(!f->gp_row_scale || f->At.row_vals, "gcn_vjp_phase2: gp_row_scale needs the unit pattern of A_hat^T in At");
  GcnWs w;
  rc = carve(f, ws, ws_bytes, w);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  const int d = f->d;
  const int64_t n = f->A.n_rows;
  float* gW = gtheta;
  float* gb = gtheta + static_cast<size_t>(d + 1) * d;
  float* ggamma = gb + d;
  float* gbeta = ggamma + d;
  float* gt = gbeta + d;
  float* gS = w.bufB;
  float* z = w.bufC;
  float* gz = w.bufA;
  float* cs = w.small;
  // gS = A_hat^T gP
  gode_spmm_epilogue_t ep;
  memset(&ep, 0, sizeof(ep));
  {
    ProfScope prof(GODE_PROF_AGG_T, st);
    ep.acc_in = f->partial_in;
    rc = spmm_dispatch(f->At, gP + f->gather_row_offset * d, d, d, gS, d, ep, w.heavy, w.heavy_bytes, st);
  }
  if (rc) return rc;
  ProfScope prof_dense(GODE_PROF_VJP_DENSE, st);
  // bias gradient: column sums of gP -- or, with gp_row_scale (gP holds row-scaled rows), of gS (below)
  if (!f->gp_row_scale && (rc = colsum(n, d, gP, d, gb, w.red, w.red_bytes, st))) return rc;
  // gW[1:,:] = z^T gS, and cs = column sums of gS (tensor-core path: same pass over gS)
  if (wgrad_tc_supported(f)) {
    if ((rc = wgrad_tc(f, y, gS, cs, gW + d, w.splitk, w.splitk_bytes, st))) return rc;
  } else {
    if ((rc = colsum(n, d, gS, d, cs, w.red, w.red_bytes, st))) return rc;
    if ((rc = groupnorm_fwd(n, d, f->groups, f->gn_eps, y, d, f->gamma, f->beta, z, d, st))) return rc;
    if ((rc = gemm_simt(1, 0, d, d, n, 1.f, z, d, gS, d, 0.f, gW + d, d, splits_for(n), w.splitk, w.splitk_bytes, st, nullptr, 0.f)))
      return rc;
  }
  if (f->gp_row_scale) GODE_CHECK_CUDA(cudaMemcpyAsync(gb, cs, sizeof(float) * d, cudaMemcpyDeviceToDevice, st));
  // the two time-column terms: gW[0,:] = t cs, gt = <cs, W[0,:]>
  k_time_terms<<<1, 128, 0, st>>>(d, t, cs, f->W, gW, gt);
  GODE_LAUNCH_CHECK();
  // gz = gS W[1:,:]^T ; GroupNorm backward with the adjoint state's Runge-Kutta combination in its tail
  if (input_grad_tc_supported(f)) {
    if ((rc = input_grad_tc(f, gS, gz, st))) return rc;
  } else if ((rc = gemm_simt(0, 1, n, d, d, 1.f, gS, d, f->W + d, d, 0.f, gz, d, 1, nullptr, 0, st, nullptr, 0.f))) {
    return rc;
  }
  return groupnorm_bwd_rk(n, d, f->groups, f->gn_eps, y, d, f->gamma, gz, d, k_a, d, ggamma, gbeta, w.red, w.red_bytes, st,
                          a0, kprev_host, coef_host, n_prev, coef_self, a_next, f->second.out ? &f->second : nullptr);
}

extern "C" int gode_gcn_vjp_phase2(const gode_gcn_odefunc_t* f, const float* y, float t, const float* gP, float* k_a,
                                   float* gtheta, void* ws, size_t ws_bytes, void* stream) {
  GODE_REQUIRE(k_a != nullptr || (f && f->A.n_rows == 0), "gcn_vjp_phase2: null pointer");
  return gode_gcn_vjp_phase2_rk(f, y, t, gP, k_a, gtheta, nullptr, nullptr, nullptr, 0, 0.f, nullptr, ws, ws_bytes, stream);
}

extern "C" int gode_gcn_stage_vjp(const gode_gcn_odefunc_t* f, const float* y, float t, const float* S, const float* a,
                                  float sign, float* k_y, float* k_a, float* gtheta, void* ws, size_t ws_bytes,
                                  void* stream) {
  int rc = check(f);
  if (rc) return rc;
  GODE_REQUIRE(f->A.n_cols == f->A.n_rows && f->At.n_cols == f->A.n_rows && f->gather_row_offset == 0 && !f->partial_in,
               "gcn_stage_vjp: partitioned graphs must call phase1 / exchange / phase2");
  GODE_REQUIRE(!f->second.out, "gcn_stage_vjp: a second Runge-Kutta combination needs the phase1 / phase2 calls");
  GcnWs w;
  rc = carve(f, ws, ws_bytes, w);
  if (rc) return rc;
  if ((rc = gode_gcn_vjp_phase1(f, S, a, sign, k_y, w.bufA, nullptr, nullptr, nullptr, 0, 0.f, nullptr, ws, ws_bytes, stream)))
    return rc;
  // phase2 reuses bufA for gz only after the bias column sum has consumed gP
  return gode_gcn_vjp_phase2(f, y, t, w.bufA, k_a, gtheta, ws, ws_bytes, stream);
}
