// libgode: dense fp32 GEMM on the SIMT pipes, any shape / transposition, deterministic split-K.
//
// Reference call sites: torch.mm(input, weight) GCN/layers.py:32,70 and the two transposed products its
// autograd issues.  This is the shape-agnostic path (input layers F x h with F = 1433/3703/500, class
// layers h x 7, QC hidden 73, weight gradients).  The square d x d products inside the ODE function at
// d = 64/128 go to the tcgen05 kernels in transform_tc.cu.
#include "internal.cuh"

namespace gode {

template <bool TA, bool TB>
struct Ld {
  __device__ static __forceinline__ float a(const float* __restrict__ A, int64_t lda, int64_t m, int64_t k) {
    return TA ? __ldg(A + k * lda + m) : __ldg(A + m * lda + k);
  }
  __device__ static __forceinline__ float b(const float* __restrict__ B, int64_t ldb, int64_t k, int64_t n) {
    return TB ? __ldg(B + n * ldb + k) : __ldg(B + k * ldb + n);
  }
};

// C tile BM x BN per CTA (256 threads as 16 x 16, TM x TN micro-tile each), K step 8.
template <int BM, int BN, bool TA, bool TB>
__global__ void __launch_bounds__(256) k_gemm(int64_t M, int64_t N, int64_t K, float alpha, const float* __restrict__ A,
                                              int64_t lda, const float* __restrict__ B, int64_t ldb, float beta,
                                              float* __restrict__ C, int64_t ldc, int64_t k_chunk, float* __restrict__ part,
                                              const float* __restrict__ rowvec, float rowvec_scale, int relu) {
  constexpr int BK = 8, TM = BM / 16, TN = BN / 16;
  __shared__ float sA[BK][BM + 4];
  __shared__ float sB[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = blockIdx.x * (int64_t)BM, n0 = blockIdx.y * (int64_t)BN;  // M tiles on x (2^31-1 limit)
  const int64_t kbeg = blockIdx.z * k_chunk;
  const int64_t kend = min(K, kbeg + k_chunk);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    // A tile: BM*BK elements; consecutive threads walk the contiguous dimension of the source
#pragma unroll
    for (int i = tid; i < BM * BK; i += 256) {
      int mm, kk;
      if (TA) { mm = i % BM; kk = i / BM; } else { kk = i % BK; mm = i / BK; }
      const int64_t m = m0 + mm, k = k0 + kk;
      sA[kk][mm] = (m < M && k < kend) ? Ld<TA, TB>::a(A, lda, m, k) : 0.f;
    }
#pragma unroll
    for (int i = tid; i < BN * BK; i += 256) {
      int nn, kk;
      if (TB) { kk = i % BK; nn = i / BK; } else { nn = i % BN; kk = i / BN; }
      const int64_t n = n0 + nn, k = k0 + kk;
      sB[kk][nn] = (n < N && k < kend) ? Ld<TA, TB>::b(B, ldb, k, n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = sA[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = sB[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + tx * TN + j;
      if (n >= N) continue;
      if (part) {
        part[(blockIdx.z * M + m) * N + n] = alpha * acc[i][j];
      } else {
        float prev = beta != 0.f ? beta * C[m * ldc + n] : 0.f;
        if (rowvec) prev += rowvec_scale * __ldg(rowvec + n);
        const float v = alpha * acc[i][j] + prev;
        C[m * ldc + n] = relu ? fmaxf(v, 0.f) : v;
      }
    }
  }
}

__global__ void k_reduce_splits(int64_t M, int64_t N, int splits, const float* __restrict__ part, float beta,
                                float* __restrict__ C, int64_t ldc) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[z * M * N + i];
  const int64_t m = i / N, n = i % N;
  C[m * ldc + n] = s + (beta != 0.f ? beta * C[m * ldc + n] : 0.f);
}

template <int BM, int BN>
static int launch_gemm(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
                       const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int splits, int64_t k_chunk,
                       float* part, const float* rowvec, float rowvec_scale, int relu, cudaStream_t st) {
  dim3 grid(static_cast<unsigned>((M + BM - 1) / BM), static_cast<unsigned>((N + BN - 1) / BN), splits);
  if (!ta && !tb) k_gemm<BM, BN, false, false><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, k_chunk, part, rowvec, rowvec_scale, relu);
  else if (!ta && tb) k_gemm<BM, BN, false, true><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, k_chunk, part, rowvec, rowvec_scale, relu);
  else if (ta && !tb) k_gemm<BM, BN, true, false><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, k_chunk, part, rowvec, rowvec_scale, relu);
  else k_gemm<BM, BN, true, true><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, k_chunk, part, rowvec, rowvec_scale, relu);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

int gemm_simt(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, const float* B,
              int64_t ldb, float beta, float* C, int64_t ldc, int splits, void* ws, size_t ws_bytes, cudaStream_t st,
              const float* rowvec, float rowvec_scale, int relu) {
  if (M == 0 || N == 0) return GODE_OK;
  if ((rowvec || relu) && splits > 1) {
    set_error("gemm: rowvec epilogue is not available with split-K");
    return GODE_EINVAL;
  }
  if (splits < 1) splits = 1;
  if (splits > K) splits = K > 0 ? static_cast<int>(K) : 1;
  float* part = nullptr;
  int64_t k_chunk = K;
  if (splits > 1) {
    if (ws_bytes < sizeof(float) * splits * M * N || !ws) {
      set_error("gemm: split-K workspace too small (%zu < %zu)", ws_bytes, sizeof(float) * (size_t)(splits * M * N));
      return GODE_EWORKSPACE;
    }
    part = static_cast<float*>(ws);
    k_chunk = ((K + splits - 1) / splits + 7) / 8 * 8;
    splits = static_cast<int>((K + k_chunk - 1) / k_chunk);
  }
  int rc;
  if (M >= 128 && N >= 128)
    rc = launch_gemm<128, 128>(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, splits, k_chunk, part, rowvec, rowvec_scale, relu, st);
  else if (N <= 32)
    rc = launch_gemm<128, 32>(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, splits, k_chunk, part, rowvec, rowvec_scale, relu, st);
  else
    rc = launch_gemm<64, 64>(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, splits, k_chunk, part, rowvec, rowvec_scale, relu, st);
  if (rc != GODE_OK) return rc;
  if (part) {
    k_reduce_splits<<<static_cast<unsigned>((M * N + 255) / 256), 256, 0, st>>>(M, N, splits, part, beta, C, ldc);
    GODE_LAUNCH_CHECK();
  }
  return GODE_OK;
}

}  // namespace gode

extern "C" int gode_gemm_f32(int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K, float alpha, const float* A,
                             int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc,
                             int32_t precision, int32_t splits, void* ws, size_t ws_bytes, void* stream) {
  using namespace gode;
  (void)precision;  // the SIMT path is always full fp32
  GODE_REQUIRE(M >= 0 && N >= 0 && K >= 0, "gemm: negative size");
  GODE_REQUIRE((M == 0 || N == 0) || (A && B && C) || K == 0, "gemm: null pointer");
  GODE_REQUIRE(ldc >= N, "gemm: ldc < N");
  return gemm_simt(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, splits, ws, ws_bytes, as_stream(stream), nullptr, 0.f);
}

// C = act(A B + bias): nn.Linear / MyLinear with the bias (and ReLU) applied in the GEMM epilogue
extern "C" int gode_linear_f32(int32_t transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                               int64_t ldb, const float* bias, int32_t relu, float* C, int64_t ldc, void* stream) {
  using namespace gode;
  GODE_REQUIRE(M >= 0 && N >= 0 && K >= 0, "linear: negative size");
  GODE_REQUIRE((M == 0 || N == 0) || (A && B && C) || K == 0, "linear: null pointer");
  GODE_REQUIRE(ldc >= N, "linear: ldc < N");
  return gemm_simt(0, transB, M, N, K, 1.f, A, lda, B, ldb, 0.f, C, ldc, 1, nullptr, 0, as_stream(stream), bias, 1.f,
                   relu ? 1 : 0);
}
