// libgode: QC edge-conditioned message kernels -- the per-edge mat-vec of the edge-conditioned convolution.
//
// Reference: EdgeGraphConvolution.forward QC/layers.py:142-147 and MPNN_enn_edge.forward QC/mpnn.py:27-29:
//     edge_support = index_select(support, 0, Esrc)                  [E, f]
//     edge_msg     = bmm(edge_data, edge_support.unsqueeze(-1))       [E, f]     edge_data [E, f, f]
//     output       = spmm(Etgt, edge_msg)                             [N, f]     Etgt dense one-hot [N, E]
//
// The bmm is HBM-bound on edge_data (f*f*4 bytes per edge: 21.3 KB at the reference's hidden = 73, about
// 0.5 flop/byte), so it is a streaming kernel, not a tensor-core one: one warp per edge reads the edge matrix
// row by row with coalesced loads (f = 73 rows are only 4-byte aligned -> scalar lanes, 128 B per instruction),
// the gathered support row sits in registers.  The dense one-hot product with Etgt -- O(N*E*f) in the
// reference -- is a segmented sum by target done by gode_spmm_csr_f32 over the [E, f] message buffer
// (graph-odenet_b200/ops.py::EdgeGraph), deterministic and without atomics.
//
// Backward of the mat-vec (given dm_e = gout[tgt_e]):  ds_e = edge_data[e]^T dm_e  and
// d edge_data[e] = dm_e (x) support[src_e]  (an [E, f, f] output: the edge encoder MLP needs it); one pass
// reads edge_data once and writes d edge_data once, lanes own columns so there is no cross-lane reduction.
#include "internal.cuh"

namespace gode {

// msg[e, r] = sum_c ed[e, r, c] * s[src[e], c]
template <int NCH>
__global__ void __launch_bounds__(256) k_edge_matvec(int64_t n_edges, int f, const float* __restrict__ ed,
                                                     const int32_t* __restrict__ src, const float* __restrict__ s, int64_t lds,
                                                     float* __restrict__ msg) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t e = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); e < n_edges; e += nwarps) {
    const float* srow = s + (int64_t)__ldg(src + e) * lds;
    float sv[NCH];
#pragma unroll
    for (int q = 0; q < NCH; ++q) {
      const int c = lane + 32 * q;
      sv[q] = c < f ? __ldg(srow + c) : 0.f;
    }
    const float* m = ed + e * (int64_t)f * f;
    for (int r0 = 0; r0 < f; r0 += 4) {       // 4 rows at a time: 4*NCH independent 128 B loads in flight per lane
      float p[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        p[u] = 0.f;
        const int r = r0 + u;
        if (r < f) {
          const float* row = m + (int64_t)r * f;
#pragma unroll
          for (int q = 0; q < NCH; ++q) {
            const int c = lane + 32 * q;
            if (c < f) p[u] += __ldcs(row + c) * sv[q];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float t = warp_sum(p[u]);
        if (lane == u && r0 + u < f) msg[e * f + r0 + u] = t;
      }
    }
  }
}

// ds[e, c] = sum_r ed[e, r, c] * g[tgt[e], r];  d_ed[e, r, c] = g[tgt[e], r] * s[src[e], c]   (d_ed may be NULL)
template <int NCH>
__global__ void __launch_bounds__(256) k_edge_matvec_bwd(int64_t n_edges, int f, const float* __restrict__ ed,
                                                         const int32_t* __restrict__ src, const int32_t* __restrict__ tgt,
                                                         const float* __restrict__ s, int64_t lds, const float* __restrict__ g,
                                                         int64_t ldg, float* __restrict__ ds, float* __restrict__ d_ed) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t e = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); e < n_edges; e += nwarps) {
    const float* srow = s + (int64_t)__ldg(src + e) * lds;
    const float* grow = g + (tgt ? (int64_t)__ldg(tgt + e) : e) * ldg;   // tgt == NULL: g is already per edge
    float sv[NCH], gv[NCH], acc[NCH];
#pragma unroll
    for (int q = 0; q < NCH; ++q) {
      const int c = lane + 32 * q;
      sv[q] = c < f ? __ldg(srow + c) : 0.f;
      gv[q] = c < f ? __ldg(grow + c) : 0.f;
      acc[q] = 0.f;
    }
    const float* m = ed + e * (int64_t)f * f;
    float* dm = d_ed ? d_ed + e * (int64_t)f * f : nullptr;
#pragma unroll
    for (int qr = 0; qr < NCH; ++qr) {
      const int rbase = 32 * qr;
      const int rcnt = min(32, f - rbase);
      for (int rr = 0; rr < rcnt; ++rr) {
        const float gr = __shfl_sync(0xffffffffu, gv[qr], rr);
        const int64_t off = (int64_t)(rbase + rr) * f;
#pragma unroll
        for (int q = 0; q < NCH; ++q) {
          const int c = lane + 32 * q;
          if (c < f) {
            acc[q] += __ldcs(m + off + c) * gr;
            if (dm) __stcs(dm + off + c, gr * sv[q]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < NCH; ++q) {
      const int c = lane + 32 * q;
      if (c < f) ds[e * f + c] = acc[q];
    }
  }
}

static unsigned edge_grid(int64_t n_edges) {
  int64_t blocks = (n_edges + 7) / 8;
  const int64_t cap = 8LL * sm_count();
  if (blocks > cap) blocks = cap;
  return static_cast<unsigned>(blocks < 1 ? 1 : blocks);
}

}  // namespace gode

using namespace gode;

extern "C" int gode_edge_matvec(int64_t n_edges, int32_t f, const float* edge_data, const int32_t* esrc, const float* s,
                                int64_t lds, float* msg, void* stream) {
  GODE_REQUIRE(n_edges >= 0 && f >= 1 && f <= 128 && lds >= f, "edge_matvec: feature width must be in [1,128]");
  if (n_edges == 0) return GODE_OK;
  GODE_REQUIRE(edge_data && esrc && s && msg, "edge_matvec: null pointer");
  cudaStream_t st = as_stream(stream);
  const unsigned grid = edge_grid(n_edges);
  switch ((f + 31) / 32) {
    case 1: k_edge_matvec<1><<<grid, 256, 0, st>>>(n_edges, f, edge_data, esrc, s, lds, msg); break;
    case 2: k_edge_matvec<2><<<grid, 256, 0, st>>>(n_edges, f, edge_data, esrc, s, lds, msg); break;
    case 3: k_edge_matvec<3><<<grid, 256, 0, st>>>(n_edges, f, edge_data, esrc, s, lds, msg); break;
    default: k_edge_matvec<4><<<grid, 256, 0, st>>>(n_edges, f, edge_data, esrc, s, lds, msg); break;
  }
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

extern "C" int gode_edge_matvec_bwd(int64_t n_edges, int32_t f, const float* edge_data, const int32_t* esrc,
                                    const int32_t* etgt, const float* s, int64_t lds, const float* gout, int64_t ldg,
                                    float* ds_msg, float* d_edge_data, void* stream) {
  GODE_REQUIRE(n_edges >= 0 && f >= 1 && f <= 128 && lds >= f && ldg >= f, "edge_matvec_bwd: feature width must be in [1,128]");
  if (n_edges == 0) return GODE_OK;
  GODE_REQUIRE(edge_data && esrc && s && gout && ds_msg, "edge_matvec_bwd: null pointer");
  cudaStream_t st = as_stream(stream);
  const unsigned grid = edge_grid(n_edges);
  switch ((f + 31) / 32) {
    case 1: k_edge_matvec_bwd<1><<<grid, 256, 0, st>>>(n_edges, f, edge_data, esrc, etgt, s, lds, gout, ldg, ds_msg, d_edge_data); break;
    case 2: k_edge_matvec_bwd<2><<<grid, 256, 0, st>>>(n_edges, f, edge_data, esrc, etgt, s, lds, gout, ldg, ds_msg, d_edge_data); break;
    case 3: k_edge_matvec_bwd<3><<<grid, 256, 0, st>>>(n_edges, f, edge_data, esrc, etgt, s, lds, gout, ldg, ds_msg, d_edge_data); break;
    default: k_edge_matvec_bwd<4><<<grid, 256, 0, st>>>(n_edges, f, edge_data, esrc, etgt, s, lds, gout, ldg, ds_msg, d_edge_data); break;
  }
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}
