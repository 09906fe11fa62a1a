// libgode: graph plan -- canonical CSR / CSR^T from the reference's COO adjacency.
//
// Reference: sparse_mx_to_torch_sparse_tensor (GCN/utils.py:222-229) hands torch.spmm an uncoalesced
// int64 COO; ATen converts it on every call.  Here the conversion happens once, on the device, and the
// result is the bit-exact canonical form defined by oracle/graph_ops.py:coo_to_csr / csr_transpose.
// One-off work (not on the per-evaluation path): the radix sort and scans come from CUB.
#include "common.cuh"
#include <cub/cub.cuh>

namespace gode {

__global__ void k_make_keys(int64_t nnz, const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                            uint64_t* __restrict__ keys, int32_t* __restrict__ idx, int64_t n_rows, int64_t n_cols,
                            int* __restrict__ bad) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  int64_t r = row[i], c = col[i];
  if (r < 0 || r >= n_rows || c < 0 || c >= n_cols) {
    *bad = 1;
    r = 0;
    c = 0;
  }
  keys[i] = (static_cast<uint64_t>(r) << 32) | static_cast<uint64_t>(c);
  idx[i] = static_cast<int32_t>(i);
}

__global__ void k_head_flags(int64_t nnz, const uint64_t* __restrict__ keys, int32_t* __restrict__ head) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// one thread per sorted entry that starts a (row,col) group: sums the group in input order
__global__ void k_merge(int64_t nnz, const uint64_t* __restrict__ keys, const int32_t* __restrict__ idx,
                        const int32_t* __restrict__ head, const int32_t* __restrict__ seg /*inclusive scan of head*/,
                        const float* __restrict__ val, int32_t* __restrict__ colidx, float* __restrict__ vals,
                        uint64_t* __restrict__ ukeys) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nnz || !head[i]) return;
  float acc = val[idx[i]];
  for (int64_t j = i + 1; j < nnz && !head[j]; ++j) acc = acc + val[idx[j]];
  int32_t s = seg[i] - 1;
  colidx[s] = static_cast<int32_t>(keys[i] & 0xffffffffull);
  vals[s] = acc;
  ukeys[s] = keys[i];
}

// rowptr[r] = number of unique keys with row < r  (lower bound of r<<32)
__global__ void k_rowptr(int64_t n_rows, const uint64_t* __restrict__ ukeys, const int32_t* __restrict__ seg,
                         int64_t nnz, int32_t* __restrict__ rowptr, int64_t* __restrict__ nnz_out) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  int64_t n_u = nnz > 0 ? seg[nnz - 1] : 0;
  uint64_t key = static_cast<uint64_t>(r) << 32;
  int64_t lo = 0, hi = n_u;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (ukeys[mid] < key) lo = mid + 1; else hi = mid;
  }
  rowptr[r] = static_cast<int32_t>(lo);
  if (r == 0 && nnz_out) *nnz_out = n_u;
}

__global__ void k_keys_from_csr_t(int64_t n_rows, int64_t nnz, const int32_t* __restrict__ rowptr,
                                  const int32_t* __restrict__ colidx, uint64_t* __restrict__ keys,
                                  int32_t* __restrict__ idx) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  // row of entry e: largest r with rowptr[r] <= e
  int64_t lo = 0, hi = n_rows;
  while (lo < hi) {
    int64_t mid = (lo + hi + 1) >> 1;
    if (rowptr[mid] <= e) lo = mid; else hi = mid - 1;
  }
  keys[e] = (static_cast<uint64_t>(static_cast<uint32_t>(colidx[e])) << 32) | static_cast<uint64_t>(lo);
  idx[e] = static_cast<int32_t>(e);
}

__global__ void k_emit_t(int64_t nnz, const uint64_t* __restrict__ keys, const int32_t* __restrict__ perm,
                         const float* __restrict__ vals, int32_t* __restrict__ colidx_t, float* __restrict__ vals_t,
                         int32_t* __restrict__ perm_t) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  colidx_t[e] = static_cast<int32_t>(keys[e] & 0xffffffffull);
  vals_t[e] = vals[perm[e]];
  if (perm_t) perm_t[e] = perm[e];
}

__global__ void k_rowptr_plain(int64_t n_rows, const uint64_t* __restrict__ keys, int64_t nnz,
                               int32_t* __restrict__ rowptr) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  uint64_t key = static_cast<uint64_t>(r) << 32;
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < key) lo = mid + 1; else hi = mid;
  }
  rowptr[r] = static_cast<int32_t>(lo);
}

__global__ void k_heavy_flags(int64_t n_rows, const int32_t* __restrict__ rowptr, int32_t* __restrict__ flag) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  flag[r] = (rowptr[r + 1] - rowptr[r] > GODE_HEAVY_ROW) ? 1 : 0;
}

__global__ void k_set_one(int32_t* p) { *p = 1; }

__global__ void k_flag_bad(const int* bad, int64_t* nnz_out) {
  if (*bad && nnz_out) *nnz_out = -1;
}

static int bits_for(int64_t n) {
  int b = 1;
  while ((int64_t(1) << b) < n) ++b;
  return b;
}

static size_t sort_temp_bytes(int64_t nnz) {
  size_t t = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, nnz);
  size_t s = 0;
  cub::DeviceScan::InclusiveSum(nullptr, s, (const int32_t*)nullptr, (int32_t*)nullptr, nnz);
  return align_up(t > s ? t : s, 256);
}

}  // namespace gode

using namespace gode;

extern "C" size_t gode_csr_from_coo_workspace_bytes(int64_t nnz, int64_t n_rows) {
  (void)n_rows;
  size_t n = static_cast<size_t>(nnz > 0 ? nnz : 1);
  // keys x2, ukeys, idx x2, head, seg, flag + cub temp
  return 3 * align_up(n * 8, 256) + 4 * align_up(n * 4, 256) + 256 + sort_temp_bytes(nnz > 0 ? nnz : 1);
}

extern "C" int gode_csr_from_coo(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* row, const int64_t* col,
                                 const float* val, int32_t* rowptr, int32_t* colidx, float* vals, int64_t* nnz_out,
                                 void* ws, size_t ws_bytes, void* stream_) {
  GODE_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0, "csr_from_coo: negative size");
  GODE_REQUIRE(n_rows < (int64_t(1) << 31) && n_cols < (int64_t(1) << 31) && nnz < (int64_t(1) << 31),
               "csr_from_coo: sizes must fit int32 (n_rows=%lld n_cols=%lld nnz=%lld)", (long long)n_rows,
               (long long)n_cols, (long long)nnz);
  GODE_REQUIRE(rowptr && (nnz == 0 || (row && col && val && colidx && vals)), "csr_from_coo: null pointer");
  if (ws_bytes < gode_csr_from_coo_workspace_bytes(nnz, n_rows)) {
    set_error("csr_from_coo: workspace too small");
    return GODE_EWORKSPACE;
  }
  cudaStream_t st = as_stream(stream_);
  const int T = 256;
  if (nnz == 0) {
    GODE_CHECK_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int32_t) * (n_rows + 1), st));
    if (nnz_out) GODE_CHECK_CUDA(cudaMemsetAsync(nnz_out, 0, sizeof(int64_t), st));
    return GODE_OK;
  }
  Arena ar(ws, ws_bytes);
  uint64_t* keys_a = ar.take<uint64_t>(nnz);
  uint64_t* keys_b = ar.take<uint64_t>(nnz);
  uint64_t* ukeys = ar.take<uint64_t>(nnz);
  int32_t* idx_a = ar.take<int32_t>(nnz);
  int32_t* idx_b = ar.take<int32_t>(nnz);
  int32_t* head = ar.take<int32_t>(nnz);
  int32_t* seg = ar.take<int32_t>(nnz);
  int* bad = ar.take<int>(1);
  size_t temp_bytes = sort_temp_bytes(nnz);
  void* temp = ar.take<char>(temp_bytes);
  GODE_REQUIRE(temp != nullptr, "csr_from_coo: arena exhausted");
  unsigned grid = static_cast<unsigned>((nnz + T - 1) / T);
  GODE_CHECK_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
  k_make_keys<<<grid, T, 0, st>>>(nnz, row, col, keys_a, idx_a, n_rows, n_cols, bad);
  GODE_LAUNCH_CHECK();
  int end_bit = 32 + bits_for(n_rows);
  if (end_bit > 64) end_bit = 64;
  GODE_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_a, keys_b, idx_a, idx_b, nnz, 0, end_bit, st));
  k_head_flags<<<grid, T, 0, st>>>(nnz, keys_b, head);
  GODE_LAUNCH_CHECK();
  GODE_CHECK_CUDA(cub::DeviceScan::InclusiveSum(temp, temp_bytes, head, seg, nnz, st));
  k_merge<<<grid, T, 0, st>>>(nnz, keys_b, idx_b, head, seg, val, colidx, vals, ukeys);
  GODE_LAUNCH_CHECK();
  unsigned grid_r = static_cast<unsigned>((n_rows + 1 + T - 1) / T);
  k_rowptr<<<grid_r, T, 0, st>>>(n_rows, ukeys, seg, nnz, rowptr, nnz_out);
  GODE_LAUNCH_CHECK();
  // out-of-range indices are reported through nnz_out = -1
  // (checked by the host binding after it reads nnz_out back)
  k_flag_bad<<<1, 1, 0, st>>>(bad, nnz_out);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

extern "C" size_t gode_csr_transpose_workspace_bytes(int64_t nnz, int64_t n_rows, int64_t n_cols) {
  (void)n_rows;
  (void)n_cols;
  size_t n = static_cast<size_t>(nnz > 0 ? nnz : 1);
  return 2 * align_up(n * 8, 256) + 2 * align_up(n * 4, 256) + sort_temp_bytes(nnz > 0 ? nnz : 1);
}

extern "C" int gode_csr_transpose(int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* rowptr,
                                  const int32_t* colidx, const float* vals, int32_t* rowptr_t, int32_t* colidx_t,
                                  float* vals_t, int32_t* perm_t, void* ws, size_t ws_bytes, void* stream_) {
  GODE_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0 && nnz < (int64_t(1) << 31), "csr_transpose: bad size");
  GODE_REQUIRE(rowptr && rowptr_t && (nnz == 0 || (colidx && vals && colidx_t && vals_t)), "csr_transpose: null pointer");
  if (ws_bytes < gode_csr_transpose_workspace_bytes(nnz, n_rows, n_cols)) {
    set_error("csr_transpose: workspace too small");
    return GODE_EWORKSPACE;
  }
  cudaStream_t st = as_stream(stream_);
  const int T = 256;
  if (nnz == 0) {
    GODE_CHECK_CUDA(cudaMemsetAsync(rowptr_t, 0, sizeof(int32_t) * (n_cols + 1), st));
    return GODE_OK;
  }
  Arena ar(ws, ws_bytes);
  uint64_t* keys_a = ar.take<uint64_t>(nnz);
  uint64_t* keys_b = ar.take<uint64_t>(nnz);
  int32_t* idx_a = ar.take<int32_t>(nnz);
  int32_t* idx_b = ar.take<int32_t>(nnz);
  size_t temp_bytes = sort_temp_bytes(nnz);
  void* temp = ar.take<char>(temp_bytes);
  GODE_REQUIRE(temp != nullptr, "csr_transpose: arena exhausted");
  unsigned grid = static_cast<unsigned>((nnz + T - 1) / T);
  k_keys_from_csr_t<<<grid, T, 0, st>>>(n_rows, nnz, rowptr, colidx, keys_a, idx_a);
  GODE_LAUNCH_CHECK();
  int end_bit = 32 + bits_for(n_cols);
  if (end_bit > 64) end_bit = 64;
  GODE_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_a, keys_b, idx_a, idx_b, nnz, 0, end_bit, st));
  k_emit_t<<<grid, T, 0, st>>>(nnz, keys_b, idx_b, vals, colidx_t, vals_t, perm_t);
  GODE_LAUNCH_CHECK();
  unsigned grid_r = static_cast<unsigned>((n_cols + 1 + T - 1) / T);
  k_rowptr_plain<<<grid_r, T, 0, st>>>(n_cols, keys_b, nnz, rowptr_t);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

namespace gode {
// single block, ordered compaction of the heavy rows + exclusive prefix of their chunk counts
// (hubs are few; one-off plan work)
__global__ void k_compact_heavy(int64_t n_rows, const int32_t* __restrict__ rowptr, int32_t* __restrict__ out,
                                int32_t* __restrict__ chunk_ptr, int32_t* __restrict__ counts) {
  __shared__ int32_t base;
  __shared__ int32_t warp_tot[32];
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int64_t start = 0; start < n_rows; start += blockDim.x) {
    int64_t r = start + threadIdx.x;
    int flag = (r < n_rows && rowptr[r + 1] - rowptr[r] > GODE_HEAVY_ROW) ? 1 : 0;
    unsigned m = __ballot_sync(0xffffffffu, flag);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int pre = __popc(m & ((1u << lane) - 1));
    if (lane == 0) warp_tot[w] = __popc(m);
    __syncthreads();
    int off = 0;
    for (int i = 0; i < w; ++i) off += warp_tot[i];
    if (flag) out[base + off + pre] = static_cast<int32_t>(r);
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int i = 0; i < (blockDim.x >> 5); ++i) tot += warp_tot[i];
      base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int32_t acc = 0;
    for (int i = 0; i < base; ++i) {
      chunk_ptr[i] = acc;
      const int32_t len = rowptr[out[i] + 1] - rowptr[out[i]];
      acc += (len + GODE_HEAVY_CHUNK - 1) / GODE_HEAVY_CHUNK;
    }
    chunk_ptr[base] = acc;
    counts[0] = base;
    counts[1] = acc;
  }
}
}  // namespace gode

extern "C" int gode_csr_heavy_rows(int64_t n_rows, const int32_t* rowptr, int32_t* heavy_rows, int32_t* heavy_chunk_ptr,
                                   int32_t* counts_out, void* stream_) {
  GODE_REQUIRE(rowptr && heavy_rows && heavy_chunk_ptr && counts_out, "csr_heavy_rows: null pointer");
  k_compact_heavy<<<1, 1024, 0, as_stream(stream_)>>>(n_rows, rowptr, heavy_rows, heavy_chunk_ptr, counts_out);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

namespace gode {
__global__ void k_row_values(int64_t n_rows, const int32_t* __restrict__ rowptr, const float* __restrict__ vals,
                             float* __restrict__ out, int32_t* __restrict__ is_const) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int e0 = rowptr[row], e1 = rowptr[row + 1];
  const float v0 = e1 > e0 ? vals[e0] : 0.f;
  bool same = true;
  for (int e = e0 + lane; e < e1; e += 32) same = same && (__float_as_uint(vals[e]) == __float_as_uint(v0));
  if (!__all_sync(0xffffffffu, same) && lane == 0) *is_const = 0;
  if (lane == 0) out[row] = v0;
}
}  // namespace gode

extern "C" int gode_csr_row_values(int64_t n_rows, const int32_t* rowptr, const float* vals, float* row_vals_out,
                                   int32_t* is_const_out, void* stream_) {
  GODE_REQUIRE(rowptr && row_vals_out && is_const_out && (n_rows == 0 || vals), "csr_row_values: null pointer");
  cudaStream_t st = as_stream(stream_);
  GODE_CHECK_CUDA(cudaMemsetAsync(is_const_out, 0, sizeof(int32_t), st));
  k_set_one<<<1, 1, 0, st>>>(is_const_out);
  GODE_LAUNCH_CHECK();
  if (n_rows > 0) {
    k_row_values<<<static_cast<unsigned>((n_rows + 7) / 8), 256, 0, st>>>(n_rows, rowptr, vals, row_vals_out, is_const_out);
    GODE_LAUNCH_CHECK();
  }
  return GODE_OK;
}
