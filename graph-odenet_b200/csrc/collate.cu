// libgode: device-side collate of molecule graphs into one block-diagonal batch.
//
// Reference: QC/datasets/utils.py:153-217 (collate_g_concat_edge_data), a Python loop over the molecules of a batch run
// by the DataLoader workers: node ids are shifted by the number of atoms before the molecule, the undirected edges of a
// molecule (sorted by (src, tgt)) are emitted twice -- edge m_acc + i as (src -> tgt) and edge M + m_acc + i as
// (tgt -> src), both carrying the same feature row -- and B[j] is the position of node j's molecule in the batch.
//
// Here the dataset lives on the device as a ragged store (node_ptr / edge_ptr over all molecules, local node ids per
// undirected edge, edges of a molecule in the reference's sorted order) and a batch is the gather of `n_sel` molecule ids:
// one kernel over the output nodes and one over the output undirected edges, each element finding its molecule by binary
// search in the exclusive scans of the selected molecules' sizes.  Integer work: bit-exact against the reference collate.
// The reference stores an edge's endpoints as they are numbered inside the molecule (utils.py:196-203 do not add n_acc):
// shift = 0 reproduces that; shift = 1 adds the molecule's node offset (the block-diagonal batch the layers expect when
// molecules number their atoms from 0).
#include "internal.cuh"

namespace gode {

__device__ __forceinline__ int upper_slot(const int64_t* __restrict__ off, int n, int64_t x) {
  // largest b in [0, n) with off[b] <= x  (off is non-decreasing, off[0] = 0, x < off[n])
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= x) lo = mid;
    else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) k_collate_nodes(int64_t N, int n_sel, const int32_t* __restrict__ sel,
                                                       const int64_t* __restrict__ node_ptr, const int64_t* __restrict__ node_off,
                                                       const float* __restrict__ X_all, int n_d, int64_t* __restrict__ B,
                                                       float* __restrict__ X) {
  // one thread per (output node, feature); feature 0 also writes B
  const int64_t total = N * n_d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i / n_d;
    const int c = static_cast<int>(i - j * n_d);
    const int b = upper_slot(node_off, n_sel, j);
    const int64_t src = __ldg(node_ptr + __ldg(sel + b)) + (j - __ldg(node_off + b));
    X[i] = __ldg(X_all + src * n_d + c);
    if (c == 0) B[j] = b;
  }
}

__global__ void __launch_bounds__(256) k_collate_edges(int64_t M, int n_sel, const int32_t* __restrict__ sel,
                                                       const int64_t* __restrict__ edge_ptr, const int64_t* __restrict__ edge_off,
                                                       const int64_t* __restrict__ end_off /*nullable*/, const int32_t* __restrict__ e_src,
                                                       const int32_t* __restrict__ e_tgt, const float* __restrict__ E_all, int e_d,
                                                       float* __restrict__ E_d, int64_t* __restrict__ E_src,
                                                       int64_t* __restrict__ E_tgt) {
  const int64_t total = M * e_d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / e_d;
    const int c = static_cast<int>(i - e * e_d);
    const int b = upper_slot(edge_off, n_sel, e);
    const int64_t ge = __ldg(edge_ptr + __ldg(sel + b)) + (e - __ldg(edge_off + b));
    const float v = __ldg(E_all + ge * e_d + c);
    E_d[e * e_d + c] = v;                 // src_edge_id = m_acc + edge_id
    E_d[(M + e) * e_d + c] = v;           // tgt_edge_id = M + m_acc + edge_id
    if (c == 0) {
      const int64_t base = end_off ? __ldg(end_off + b) : 0;
      const int64_t s = base + __ldg(e_src + ge), t = base + __ldg(e_tgt + ge);
      E_src[e] = s;
      E_tgt[e] = t;
      E_src[M + e] = t;
      E_tgt[M + e] = s;
    }
  }
}

}  // namespace gode

using namespace gode;

extern "C" int gode_qc_collate(int32_t n_sel, const int32_t* sel, const int64_t* node_ptr, const int64_t* edge_ptr,
                               const int64_t* node_off, const int64_t* edge_off, int32_t shift, int64_t N, int64_t M,
                               const float* X_all, int32_t n_d, const int32_t* e_src_local, const int32_t* e_tgt_local, const float* E_all,
                               int32_t e_d, int64_t* B, float* X, float* E_d, int64_t* E_src, int64_t* E_tgt, void* stream) {
  GODE_REQUIRE(n_sel >= 0 && N >= 0 && M >= 0 && n_d >= 1 && e_d >= 1, "qc_collate: bad sizes");
  if (n_sel == 0 || (N == 0 && M == 0)) return GODE_OK;
  GODE_REQUIRE(sel && node_ptr && edge_ptr && node_off && edge_off, "qc_collate: null index arrays");
  cudaStream_t st = as_stream(stream);
  const int64_t cap = 16LL * sm_count();
  if (N > 0) {
    GODE_REQUIRE(X_all && B && X, "qc_collate: null node arrays");
    int64_t blocks = (N * n_d + 255) / 256;
    if (blocks > cap) blocks = cap;
    k_collate_nodes<<<static_cast<unsigned>(blocks), 256, 0, st>>>(N, n_sel, sel, node_ptr, node_off, X_all, n_d, B, X);
    GODE_LAUNCH_CHECK();
  }
  if (M > 0) {
    GODE_REQUIRE(e_src_local && e_tgt_local && E_all && E_d && E_src && E_tgt, "qc_collate: null edge arrays");
    int64_t blocks = (M * e_d + 255) / 256;
    if (blocks > cap) blocks = cap;
    k_collate_edges<<<static_cast<unsigned>(blocks), 256, 0, st>>>(M, n_sel, sel, edge_ptr, edge_off, shift ? node_off : nullptr, e_src_local,
                                                                   e_tgt_local, E_all, e_d, E_d, E_src, E_tgt);
    GODE_LAUNCH_CHECK();
  }
  return GODE_OK;
}
