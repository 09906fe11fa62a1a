// Internal (non-ABI) entry points shared between the translation units of libgode.
#pragma once
#include "common.cuh"

namespace gode {
size_t spmm_ws_bytes(const gode_csr_t& A, int d);
int spmm_dispatch(const gode_csr_t& A, const float* X, int64_t ldx, int32_t d, float* Y, int64_t ldy,
                  const gode_spmm_epilogue_t& ep, void* ws, size_t ws_bytes, cudaStream_t st, int64_t row_begin = 0,
                  int64_t row_end = -1 /*rows [row_begin, row_end) only: d = 128 path; -1 = all rows*/);
int gemm_simt(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, const float* B,
              int64_t ldb, float beta, float* C, int64_t ldc, int splits, void* ws, size_t ws_bytes, cudaStream_t st,
              const float* rowvec, float rowvec_scale, int relu = 0);
int colsum(int64_t n, int d, const float* x, int64_t ldx, float* out, void* ws, size_t ws_bytes, cudaStream_t st);
int groupnorm_fwd(int64_t n, int d, int groups, float eps, const float* x, int64_t ldx, const float* gamma,
                  const float* beta, float* y, int64_t ldy, cudaStream_t st);
int groupnorm_bwd(int64_t n, int d, int groups, float eps, const float* x, int64_t ldx, const float* gamma,
                  const float* dy, int64_t lddy, float* dx, int64_t lddx, float* dgamma, float* dbeta, void* ws,
                  size_t ws_bytes, cudaStream_t st);
int groupnorm_bwd_rk(int64_t n, int d, int groups, float eps, const float* x, int64_t ldx, const float* gamma,
                     const float* dy, int64_t lddy, float* dx, int64_t lddx, float* dgamma, float* dbeta, void* ws,
                     size_t ws_bytes, cudaStream_t st, const float* a0, const float* const* kprev, const float* coef,
                     int n_prev, float coef_self, float* a_next, const gode_rk_second_t* second = nullptr);
int rk_combine(int64_t n, const float* y0, const float* const* k, const float* c, int nk, float* out, cudaStream_t st);
}  // namespace gode
