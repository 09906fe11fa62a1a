// libgode: row pack / unpack for the halo exchange of the row-partitioned (multi-GPU) path.
//
// The reference has no distributed code; a row-partitioned torch.spmm(adj, support)
// (GCN/layers.py:71) needs the rows of `support` that other ranks own.  Each rank packs the rows its peers
// reference into one contiguous send buffer (ordered by destination rank, then by row id), the collective
// (NCCL all-to-all-v, issued by the host code) delivers them straight into the halo tail of the peers'
// extended operand buffers.
#include "internal.cuh"

namespace gode {

// dst[i, :] = src[idx[i], :]   -- one sub-warp of d/4 lanes per row, 128-bit accesses
__global__ void __launch_bounds__(256) k_gather_rows4(int64_t n_idx, const int32_t* __restrict__ idx, int d4,
                                                      const float4* __restrict__ src, int64_t lds4,
                                                      float4* __restrict__ dst, int64_t ldd4) {
  const int64_t total = n_idx * d4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / d4;
    const int c = static_cast<int>(i - r * d4);
    const int s = __ldg(idx + r);
    __stcs(dst + r * ldd4 + c, __ldg(src + (int64_t)s * lds4 + c));
  }
}

__global__ void __launch_bounds__(256) k_gather_rows1(int64_t n_idx, const int32_t* __restrict__ idx, int d,
                                                      const float* __restrict__ src, int64_t lds, float* __restrict__ dst,
                                                      int64_t ldd) {
  const int64_t total = n_idx * d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / d;
    const int c = static_cast<int>(i - r * d);
    dst[r * ldd + c] = __ldg(src + (int64_t)__ldg(idx + r) * lds + c);
  }
}

}  // namespace gode

using namespace gode;

extern "C" int gode_gather_rows(int64_t n_idx, const int32_t* idx, int32_t d, const float* src, int64_t lds, float* dst,
                                int64_t ldd, void* stream) {
  GODE_REQUIRE(n_idx >= 0 && d > 0 && lds >= d && ldd >= d, "gather_rows: bad shape");
  if (n_idx == 0) return GODE_OK;
  GODE_REQUIRE(idx && src && dst, "gather_rows: null pointer");
  cudaStream_t st = as_stream(stream);
  const bool v4 = (d % 4 == 0) && (lds % 4 == 0) && (ldd % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                  ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  const int64_t work = n_idx * (v4 ? d / 4 : d);
  int64_t blocks = (work + 255) / 256;
  const int64_t cap = 16LL * sm_count();
  if (blocks > cap) blocks = cap;
  if (v4)
    k_gather_rows4<<<static_cast<unsigned>(blocks), 256, 0, st>>>(n_idx, idx, d / 4, reinterpret_cast<const float4*>(src),
                                                                  lds / 4, reinterpret_cast<float4*>(dst), ldd / 4);
  else
    k_gather_rows1<<<static_cast<unsigned>(blocks), 256, 0, st>>>(n_idx, idx, d, src, lds, dst, ldd);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}
