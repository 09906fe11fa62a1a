// libgode: content-based attention readout over the nodes of each graph of a batch -- the inner step of Set2Set.
//
// Reference: QC/set2set.py:60-75 -- for every processing step
//     e_i = <x_i, q[batch_i]>;  a = softmax of e inside each graph;  r_g = sum_{i in g} a_i x_i
// which the reference evaluates with a Python loop over the graphs of the batch and boolean masks (:66-70) followed by a
// scatter_add (:73).  Here one warp owns one graph (nodes of a graph are contiguous: gptr[B+1]), lanes own channels:
// three short passes over the graph's rows (scores + max, normaliser, weighted sum), no atomics, fixed order.
//
// Backward, given dr [B, h]:  da_i = <dr_g, x_i>;  t = sum_i a_i da_i;  de_i = a_i (da_i - t)
//     dx_i = a_i dr_g + de_i q_g;   dq_g = sum_i de_i x_i
#include "internal.cuh"

namespace gode {

constexpr int RO_MAXC = 8;   // channels per lane: h <= 256

__global__ void __launch_bounds__(128) k_seg_attend_fwd(int n_graphs, int h, const int32_t* __restrict__ gptr,
                                                        const float* __restrict__ x, int64_t ldx, const float* __restrict__ q,
                                                        int64_t ldq, float* __restrict__ a, float* __restrict__ r, int64_t ldr) {
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (g >= n_graphs) return;
  const int i0 = gptr[g], i1 = gptr[g + 1];
  float qc[RO_MAXC], acc[RO_MAXC];
#pragma unroll
  for (int k = 0; k < RO_MAXC; ++k) {
    const int c = lane + 32 * k;
    qc[k] = c < h ? __ldg(q + (int64_t)g * ldq + c) : 0.f;
    acc[k] = 0.f;
  }
  float mx = -INFINITY;
  for (int i = i0; i < i1; ++i) {
    float e = 0.f;
#pragma unroll
    for (int k = 0; k < RO_MAXC; ++k) {
      const int c = lane + 32 * k;
      if (c < h) e += __ldg(x + (int64_t)i * ldx + c) * qc[k];
    }
    e = warp_sum(e);
    if (lane == 0) a[i] = e;
    mx = fmaxf(mx, e);
  }
  __syncwarp();
  float s = 0.f;
  for (int i = i0; i < i1; ++i) s += expf(a[i] - mx);      // every lane reads the same scores: same order, same sum
  const float inv = 1.f / s;
  for (int i = i0; i < i1; ++i) {
    const float w = expf(a[i] - mx) * inv;
#pragma unroll
    for (int k = 0; k < RO_MAXC; ++k) {
      const int c = lane + 32 * k;
      if (c < h) acc[k] += w * __ldg(x + (int64_t)i * ldx + c);
    }
    __syncwarp();
    if (lane == 0) a[i] = w;
  }
#pragma unroll
  for (int k = 0; k < RO_MAXC; ++k) {
    const int c = lane + 32 * k;
    if (c < h) r[(int64_t)g * ldr + c] = acc[k];
  }
}

__global__ void __launch_bounds__(128) k_seg_attend_bwd(int n_graphs, int h, const int32_t* __restrict__ gptr,
                                                        const float* __restrict__ x, int64_t ldx, const float* __restrict__ q,
                                                        int64_t ldq, const float* __restrict__ a, const float* __restrict__ dr,
                                                        int64_t lddr, float* __restrict__ dx, int64_t lddx,
                                                        float* __restrict__ dq, int64_t lddq) {
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (g >= n_graphs) return;
  const int i0 = gptr[g], i1 = gptr[g + 1];
  float qc[RO_MAXC], drc[RO_MAXC], dqc[RO_MAXC];
#pragma unroll
  for (int k = 0; k < RO_MAXC; ++k) {
    const int c = lane + 32 * k;
    qc[k] = c < h ? __ldg(q + (int64_t)g * ldq + c) : 0.f;
    drc[k] = c < h ? __ldg(dr + (int64_t)g * lddr + c) : 0.f;
    dqc[k] = 0.f;
  }
  float t = 0.f;
  for (int i = i0; i < i1; ++i) {
    float da = 0.f;
#pragma unroll
    for (int k = 0; k < RO_MAXC; ++k) {
      const int c = lane + 32 * k;
      if (c < h) da += drc[k] * __ldg(x + (int64_t)i * ldx + c);
    }
    da = warp_sum(da);
    t += __ldg(a + i) * da;
  }
  for (int i = i0; i < i1; ++i) {
    float xv[RO_MAXC];
    float da = 0.f;
#pragma unroll
    for (int k = 0; k < RO_MAXC; ++k) {
      const int c = lane + 32 * k;
      xv[k] = c < h ? __ldg(x + (int64_t)i * ldx + c) : 0.f;
      da += drc[k] * xv[k];
    }
    da = warp_sum(da);
    const float ai = __ldg(a + i);
    const float de = ai * (da - t);
#pragma unroll
    for (int k = 0; k < RO_MAXC; ++k) {
      const int c = lane + 32 * k;
      if (c < h) {
        dx[(int64_t)i * lddx + c] = ai * drc[k] + de * qc[k];
        dqc[k] += de * xv[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < RO_MAXC; ++k) {
    const int c = lane + 32 * k;
    if (c < h) dq[(int64_t)g * lddq + c] = dqc[k];
  }
}

}  // namespace gode

using namespace gode;

extern "C" int gode_segment_attend_fwd(int32_t n_graphs, int32_t h, const int32_t* gptr, const float* x, int64_t ldx,
                                       const float* q, int64_t ldq, float* a, float* r, int64_t ldr, void* stream) {
  GODE_REQUIRE(n_graphs >= 0 && h >= 1 && h <= 32 * RO_MAXC && ldx >= h && ldq >= h && ldr >= h, "segment_attend: bad shape");
  if (n_graphs == 0) return GODE_OK;
  GODE_REQUIRE(gptr && x && q && a && r, "segment_attend: null pointer");
  k_seg_attend_fwd<<<(n_graphs + 3) / 4, 128, 0, as_stream(stream)>>>(n_graphs, h, gptr, x, ldx, q, ldq, a, r, ldr);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

extern "C" int gode_segment_attend_bwd(int32_t n_graphs, int32_t h, const int32_t* gptr, const float* x, int64_t ldx,
                                       const float* q, int64_t ldq, const float* a, const float* dr, int64_t lddr, float* dx,
                                       int64_t lddx, float* dq, int64_t lddq, void* stream) {
  GODE_REQUIRE(n_graphs >= 0 && h >= 1 && h <= 32 * RO_MAXC && ldx >= h && ldq >= h && lddr >= h && lddx >= h && lddq >= h,
               "segment_attend_bwd: bad shape");
  if (n_graphs == 0) return GODE_OK;
  GODE_REQUIRE(gptr && x && q && a && dr && dx && dq, "segment_attend_bwd: null pointer");
  k_seg_attend_bwd<<<(n_graphs + 3) / 4, 128, 0, as_stream(stream)>>>(n_graphs, h, gptr, x, ldx, q, ldq, a, dr, lddr, dx, lddx,
                                                                       dq, lddq);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}
