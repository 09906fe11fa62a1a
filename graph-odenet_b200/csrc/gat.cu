// libgode: GAT attention aggregation -- segmented softmax / scatter over the edge list, forward and backward.
//
// Reference: GAT/layers.py:31-58 (GraphConvolution.forward) and :95-122 (FixedGraphConvolution.forward):
//     h = [x[src] || x[tgt]]                     [E, 2i]
//     y = relu(f(h));  a = w(h)                  f: Linear(2i -> o), w: Linear(2i -> 1)
//     a_exp = exp(a - max_over_ALL_edges(a))     (a global shift, NOT a per-target one)
//     out = (Mtgt (y * a_exp)) / (Mtgt a_exp + eps)          Mtgt[n, e] = 1 iff tgt[e] == n
//
// Restructuring (algebraically identical, rounding differs at the 1e-7 level):
//   * a Linear on a concatenation is a sum of two Linears, so the host first projects the NODES once,
//         P = x * [Wf_src^T | Wf_tgt^T | ww_src^T | ww_tgt^T]        [N, 2C + 2H]   (one GEMM, gode_gemm_f32)
//     with the biases folded into the target halves; then per edge  z_e = Ps[src] + Pt[tgt],  a_e = as[src] + at[tgt].
//     The [E, 2i] gather + cat + edge-level GEMM of the reference is never materialised (F up to 3703 on the
//     input layer: 2F*E*4 bytes = 138 MB per call on Citeseer become an N x 2C projection).
//   * the incidence products are segmented sums: edges are grouped by target (CSR of Mtgt), one thread owns one
//     (node, head) pair and walks its segment in edge order -- no atomics, deterministic, same summation order as
//     the reference's row-sorted COO product.
//   * heads: H independent reference heads batched in the channel dimension (C = H * oh channels; head h owns
//     channels [h*oh, (h+1)*oh)), each with its own global max.  H = 1 is the reference layer.
//
// Layout of P (row-major, leading dimension ldp >= 2C + 2H): [0,C) Ps | [C,2C) Pt | [2C,2C+H) as | [2C+H,2C+2H) at.
//
// Backward (given g = dL/dout), with w_e = exp(a_e - amax), y_e = relu(z_e), gn = g / den:
//     d z_e   = w_e * gn[tgt] * (z_e > 0)
//     d w_e   = sum_c y_e[c] gn[tgt, c]  -  sum_c out[tgt, c] gn[tgt, c]            (per head)
//     d a_e   = d w_e * w_e ;  d amax = - sum_e d a_e  -> added to the arg-max edge (torch.max's backward)
//     dPt[n]  = sum_{e: tgt=n} d z_e,  dat[n] = sum d a_e        (target pass, registers)
//     dPs[s]  = sum_{e: src=s} d z_e,  das[s] = sum d a_e        (source pass over the by-source grouping)
// The host then finishes with two GEMMs (dx = dP W^T, dW = x^T dP).
#include "internal.cuh"

namespace gode {

__device__ __forceinline__ unsigned int f2ord(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ float key_max(unsigned long long k) { return ord2f(static_cast<unsigned int>(k >> 32)); }
__device__ __forceinline__ unsigned int key_pos(unsigned long long k) { return 0xffffffffu - static_cast<unsigned int>(k); }

// amax_key[h] = max over edges of (ordered(a_e), first position on ties); edges in by-target order
__global__ void __launch_bounds__(256) k_gat_amax(int64_t n_edges, int H, const int32_t* __restrict__ t_src,
                                                  const int32_t* __restrict__ t_tgt, const float* __restrict__ P,
                                                  int64_t ldp, int C, unsigned long long* __restrict__ amax_key,
                                                  int* __restrict__ nan_flag) {
  __shared__ unsigned long long sm[8];
  for (int h = 0; h < H; ++h) {
    unsigned long long best = 0ull;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
      const float a = __ldg(P + (int64_t)__ldg(t_src + e) * ldp + 2 * C + h) + __ldg(P + (int64_t)__ldg(t_tgt + e) * ldp + 2 * C + H + h);
      if (a != a) *nan_flag = 1;
      const unsigned long long k = (static_cast<unsigned long long>(f2ord(a)) << 32) | (0xffffffffu - static_cast<unsigned int>(e));
      best = k > best ? k : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < (blockDim.x >> 5); ++i) best = sm[i] > best ? sm[i] : best;
      if (best) atomicMax(amax_key + h, best);
    }
    __syncthreads();
  }
}

// Segments.  One thread owns one (segment, head) pair and walks the segment in edge order.  A segment is a whole node
// (LIGHT pass: nodes with more than GODE_GAT_CHUNK edges are skipped) or one chunk of GODE_GAT_CHUNK consecutive edges
// of such a hub node (CHUNK pass: partial sums go to a workspace, a finishing kernel adds a node's chunks in chunk
// order -- deterministic, no atomics).  Without the split a power-law hub (4.7e4 in-edges on the 1 M-node
// Citeseer-shaped graph) is one thread's sequential loop: 57 ms for the source pass alone.
struct GatHeavy {
  int n_heavy, n_chunks;
  const int32_t* nodes;       // [n_heavy] hub nodes
  const int32_t* cptr;        // [n_heavy + 1] chunk range of each hub
  const int32_t* chunk_node;  // [n_chunks]
  const int32_t* chunk_e0;    // [n_chunks] first edge of the chunk
};

template <bool CHUNK>
__device__ __forceinline__ bool gat_segment(int64_t seg, const int32_t* __restrict__ ptr, const GatHeavy& hv, int64_t& n, int& e0,
                                            int& e1) {
  if (CHUNK) {
    n = __ldg(hv.chunk_node + seg);
    e0 = __ldg(hv.chunk_e0 + seg);
    e1 = min(e0 + GODE_GAT_CHUNK, __ldg(ptr + n + 1));
    return true;
  }
  n = seg;
  e0 = __ldg(ptr + n);
  e1 = __ldg(ptr + n + 1);
  return (e1 - e0) <= GODE_GAT_CHUNK;
}

// one thread per (segment, head); CH >= oh channels held in registers
template <int CH, bool CHUNK>
__global__ void __launch_bounds__(128) k_gat_fwd(int64_t n_seg, int H, int oh, const int32_t* __restrict__ tptr,
                                                 const int32_t* __restrict__ t_src, const float* __restrict__ P, int64_t ldp,
                                                 float eps, const unsigned long long* __restrict__ amax_key,
                                                 float* __restrict__ out, int64_t ldo, float* __restrict__ den_out,
                                                 int* __restrict__ nan_flag, GatHeavy hv, float* __restrict__ part) {
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (gid >= n_seg * H) return;
  const int64_t seg = gid / H;
  const int h = static_cast<int>(gid - seg * H);
  int64_t n;
  int e0, e1;
  if (!gat_segment<CHUNK>(seg, tptr, hv, n, e0, e1)) return;
  const int C = H * oh;
  const float amax = key_max(amax_key[h]);
  const float* pn = P + n * ldp;
  const float at = __ldg(pn + 2 * C + H + h);
  float pt[CH], num[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    pt[j] = j < oh ? __ldg(pn + C + h * oh + j) : 0.f;
    num[j] = 0.f;
  }
  float den = 0.f;
  for (int e = e0; e < e1; ++e) {
    const float* ps = P + (int64_t)__ldg(t_src + e) * ldp;
    const float w = expf(__ldg(ps + 2 * C + h) + at - amax);
    den += w;
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (j < oh) num[j] += fmaxf(__ldg(ps + h * oh + j) + pt[j], 0.f) * w;
  }
  if (CHUNK) {
    float* p = part + gid * (CH + 1);
#pragma unroll
    for (int j = 0; j < CH; ++j) p[j] = num[j];
    p[CH] = den;
    return;
  }
  den += eps;
  den_out[n * H + h] = den;
  bool bad = den != den;
#pragma unroll
  for (int j = 0; j < CH; ++j)
    if (j < oh) {
      const float o = num[j] / den;
      bad = bad || (o != o);
      out[n * ldo + h * oh + j] = o;
    }
  if (bad) *nan_flag = 1;
}

template <int CH>
__global__ void __launch_bounds__(128) k_gat_fwd_finish(int H, int oh, float eps, GatHeavy hv, const float* __restrict__ part,
                                                        float* __restrict__ out, int64_t ldo, float* __restrict__ den_out,
                                                        int* __restrict__ nan_flag) {
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (gid >= (int64_t)hv.n_heavy * H) return;
  const int i = static_cast<int>(gid / H), h = static_cast<int>(gid % H);
  const int64_t n = hv.nodes[i];
  float num[CH], den = 0.f;
#pragma unroll
  for (int j = 0; j < CH; ++j) num[j] = 0.f;
  for (int c = hv.cptr[i]; c < hv.cptr[i + 1]; ++c) {
    const float* p = part + ((int64_t)c * H + h) * (CH + 1);
#pragma unroll
    for (int j = 0; j < CH; ++j) num[j] += p[j];
    den += p[CH];
  }
  den += eps;
  den_out[n * H + h] = den;
  bool bad = den != den;
#pragma unroll
  for (int j = 0; j < CH; ++j)
    if (j < oh) {
      const float o = num[j] / den;
      bad = bad || (o != o);
      out[n * ldo + h * oh + j] = o;
    }
  if (bad) *nan_flag = 1;
}

// target pass of the backward: dPt, dat (registers), per-edge d a_e (by-target order) for the source pass
template <int CH, bool CHUNK>
__global__ void __launch_bounds__(128) k_gat_bwd_tgt(int64_t n_seg, int H, int oh, const int32_t* __restrict__ tptr,
                                                     const int32_t* __restrict__ t_src, const float* __restrict__ P,
                                                     int64_t ldp, const unsigned long long* __restrict__ amax_key,
                                                     const float* __restrict__ out, int64_t ldo,
                                                     const float* __restrict__ den, const float* __restrict__ g, int64_t ldg,
                                                     float* __restrict__ dP, float* __restrict__ dA, GatHeavy hv,
                                                     float* __restrict__ part) {
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (gid >= n_seg * H) return;
  const int64_t seg = gid / H;
  const int h = static_cast<int>(gid - seg * H);
  int64_t n;
  int e0, e1;
  if (!gat_segment<CHUNK>(seg, tptr, hv, n, e0, e1)) return;
  const int C = H * oh;
  const float amax = key_max(amax_key[h]);
  const float* pn = P + n * ldp;
  const float at = __ldg(pn + 2 * C + H + h);
  const float inv = 1.f / den[n * H + h];
  float pt[CH], gn[CH], dpt[CH];
  float go = 0.f;
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    pt[j] = 0.f; gn[j] = 0.f; dpt[j] = 0.f;
    if (j < oh) {
      pt[j] = __ldg(pn + C + h * oh + j);
      gn[j] = __ldg(g + n * ldg + h * oh + j) * inv;
      go += gn[j] * __ldg(out + n * ldo + h * oh + j);
    }
  }
  float dat = 0.f;
  for (int e = e0; e < e1; ++e) {
    const float* ps = P + (int64_t)__ldg(t_src + e) * ldp;
    const float w = expf(__ldg(ps + 2 * C + h) + at - amax);
    float dw = -go;
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (j < oh) {
        const float z = __ldg(ps + h * oh + j) + pt[j];
        if (z > 0.f) {
          dw += z * gn[j];
          dpt[j] += w * gn[j];
        }
      }
    const float da = dw * w;
    dat += da;
    dA[(int64_t)e * H + h] = da;
  }
  if (CHUNK) {
    float* p = part + gid * (CH + 1);
#pragma unroll
    for (int j = 0; j < CH; ++j) p[j] = dpt[j];
    p[CH] = dat;
    return;
  }
  float* dpn = dP + n * ldp;
#pragma unroll
  for (int j = 0; j < CH; ++j)
    if (j < oh) dpn[C + h * oh + j] = dpt[j];
  dpn[2 * C + H + h] = dat;
}

// source pass: dPs, das over the by-source grouping (s_tgt = target of the edge, s_pos = its by-target position)
template <int CH, bool CHUNK>
__global__ void __launch_bounds__(128) k_gat_bwd_src(int64_t n_seg, int H, int oh, const int32_t* __restrict__ sptr,
                                                     const int32_t* __restrict__ s_tgt, const int32_t* __restrict__ s_pos,
                                                     const float* __restrict__ P, int64_t ldp,
                                                     const unsigned long long* __restrict__ amax_key,
                                                     const float* __restrict__ den, const float* __restrict__ g, int64_t ldg,
                                                     const float* __restrict__ dA, float* __restrict__ dP, GatHeavy hv,
                                                     float* __restrict__ part) {
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (gid >= n_seg * H) return;
  const int64_t seg = gid / H;
  const int h = static_cast<int>(gid - seg * H);
  int64_t s;
  int e0, e1;
  if (!gat_segment<CHUNK>(seg, sptr, hv, s, e0, e1)) return;
  const int C = H * oh;
  const float amax = key_max(amax_key[h]);
  const float* psn = P + s * ldp;
  const float as = __ldg(psn + 2 * C + h);
  float ps[CH], dps[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    ps[j] = j < oh ? __ldg(psn + h * oh + j) : 0.f;
    dps[j] = 0.f;
  }
  float das = 0.f;
  for (int e = e0; e < e1; ++e) {
    const int64_t n = __ldg(s_tgt + e);
    const float* pt = P + n * ldp;
    const float w = expf(as + __ldg(pt + 2 * C + H + h) - amax) / __ldg(den + n * H + h);
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (j < oh) {
        const float z = ps[j] + __ldg(pt + C + h * oh + j);
        if (z > 0.f) dps[j] += w * __ldg(g + n * ldg + h * oh + j);
      }
    das += __ldg(dA + (int64_t)__ldg(s_pos + e) * H + h);
  }
  if (CHUNK) {
    float* p = part + gid * (CH + 1);
#pragma unroll
    for (int j = 0; j < CH; ++j) p[j] = dps[j];
    p[CH] = das;
    return;
  }
  float* dpn = dP + s * ldp;
#pragma unroll
  for (int j = 0; j < CH; ++j)
    if (j < oh) dpn[h * oh + j] = dps[j];
  dpn[2 * C + h] = das;
}

// ------------------------------------------------------------------------------------------------
// Lane-group kernels (C = H * oh a multiple of 4 up to 128, oh a power of two >= 4, 16-byte aligned rows).
// LPN = C / 4 lanes own one segment: lane j holds channels [4j, 4j + 4) of ALL heads' concatenated channels (head
// h = 4j / oh), so a row of P / out / g is ONE coalesced 128-bit access per lane -- 512 bytes per instruction at C = 128
// instead of the scalar kernels' 32 lanes x 4 bytes at a 64-byte stride (8x the L1 wavefronts; ncu round 2: 0.92 ms per
// forward at N = 1 M for 1.8 GB of compulsory traffic).  Per (node, head, channel) the edges are still walked in edge
// order by one lane: the segmented sums keep the reference's summation order; only the per-head dot products of the
// backward (over oh channels) become a fixed-shape tree over the head's lanes.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int LPN>
__device__ __forceinline__ float head_sum(float v, int lanes_per_head) {
#pragma unroll
  for (int o = 1; o < LPN; o <<= 1)
    if (o < lanes_per_head) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LPN, bool CHUNK>
__global__ void __launch_bounds__(128) k_gat_fwd_v(int64_t n_seg, int H, int oh, const int32_t* __restrict__ tptr,
                                                   const int32_t* __restrict__ t_src, const float* __restrict__ P, int64_t ldp,
                                                   float eps, const unsigned long long* __restrict__ amax_key,
                                                   float* __restrict__ out, int64_t ldo, float* __restrict__ den_out,
                                                   int* __restrict__ nan_flag, GatHeavy hv, float* __restrict__ part, int CH) {
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t seg = gid / LPN;
  const int j = static_cast<int>(gid % LPN);
  if (seg >= n_seg) return;
  int64_t n;
  int e0, e1;
  if (!gat_segment<CHUNK>(seg, tptr, hv, n, e0, e1)) return;
  const int C = H * oh;
  const int h = (4 * j) / oh;
  const float amax = key_max(amax_key[h]);
  const float* pn = P + n * ldp;
  const float at = __ldg(pn + 2 * C + H + h);
  const float4 pt = ldg4(pn + C + 4 * j);
  float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
  float den = 0.f;
  for (int e = e0; e < e1; ++e) {
    const float* ps = P + (int64_t)__ldg(t_src + e) * ldp;
    const float w = expf(__ldg(ps + 2 * C + h) + at - amax);
    const float4 v = ldg4(ps + 4 * j);
    den += w;
    num.x += fmaxf(v.x + pt.x, 0.f) * w;
    num.y += fmaxf(v.y + pt.y, 0.f) * w;
    num.z += fmaxf(v.z + pt.z, 0.f) * w;
    num.w += fmaxf(v.w + pt.w, 0.f) * w;
  }
  const bool head_lead = (4 * j) % oh == 0;
  if (CHUNK) {
    float* p = part + (seg * H + h) * (CH + 1) + (4 * j - h * oh);
    p[0] = num.x; p[1] = num.y; p[2] = num.z; p[3] = num.w;
    if (head_lead) part[(seg * H + h) * (CH + 1) + CH] = den;
    return;
  }
  den += eps;
  if (head_lead) den_out[n * H + h] = den;
  const float4 o = make_float4(num.x / den, num.y / den, num.z / den, num.w / den);
  if (den != den || o.x != o.x || o.y != o.y || o.z != o.z || o.w != o.w) *nan_flag = 1;
  *reinterpret_cast<float4*>(out + n * ldo + 4 * j) = o;
}

template <int LPN, bool CHUNK>
__global__ void __launch_bounds__(128) k_gat_bwd_tgt_v(int64_t n_seg, int H, int oh, const int32_t* __restrict__ tptr,
                                                       const int32_t* __restrict__ t_src, const float* __restrict__ P,
                                                       int64_t ldp, const unsigned long long* __restrict__ amax_key,
                                                       const float* __restrict__ out, int64_t ldo,
                                                       const float* __restrict__ den, const float* __restrict__ g, int64_t ldg,
                                                       float* __restrict__ dP, float* __restrict__ dA, GatHeavy hv,
                                                       float* __restrict__ part, int CH) {
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t seg_raw = gid / LPN;
  const int j = static_cast<int>(gid % LPN);
  // every lane of a warp takes part in the shuffles: out-of-range / skipped segments run with an empty edge range
  const bool live = seg_raw < n_seg;
  const int64_t seg = live ? seg_raw : 0;
  int64_t n = 0;
  int e0 = 0, e1 = 0;
  const bool mine = live && gat_segment<CHUNK>(seg, tptr, hv, n, e0, e1);
  if (!mine) e0 = e1 = 0;
  const int C = H * oh;
  const int h = (4 * j) / oh;
  const int lph = oh / 4;
  const float amax = key_max(amax_key[h]);
  const float* pn = P + n * ldp;
  float at = 0.f, inv = 0.f, go = 0.f;
  float4 pt = make_float4(0.f, 0.f, 0.f, 0.f), gn = pt;
  if (mine) {
    at = __ldg(pn + 2 * C + H + h);
    inv = 1.f / den[n * H + h];
    pt = ldg4(pn + C + 4 * j);
    const float4 gg = ldg4(g + n * ldg + 4 * j), oo = ldg4(out + n * ldo + 4 * j);
    gn = make_float4(gg.x * inv, gg.y * inv, gg.z * inv, gg.w * inv);
    go = gn.x * oo.x + gn.y * oo.y + gn.z * oo.z + gn.w * oo.w;
  }
  go = head_sum<LPN>(go, lph);
  // lanes of one warp walk segments of different lengths: the shuffles below need the whole warp, so the loop runs to the
  // longest range of the warp and lanes past their own range contribute zeros
  int len = e1 - e0;
  int maxlen = len;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
  float4 dpt = make_float4(0.f, 0.f, 0.f, 0.f);
  float dat = 0.f;
  const bool head_lead = (4 * j) % oh == 0;
  for (int i = 0; i < maxlen; ++i) {
    const bool on = i < len;
    float w = 0.f, part_dw = 0.f;
    if (on) {
      const float* ps = P + (int64_t)__ldg(t_src + e0 + i) * ldp;
      w = expf(__ldg(ps + 2 * C + h) + at - amax);
      const float4 v = ldg4(ps + 4 * j);
      const float zx = v.x + pt.x, zy = v.y + pt.y, zz = v.z + pt.z, zw = v.w + pt.w;
      if (zx > 0.f) { part_dw += zx * gn.x; dpt.x += w * gn.x; }
      if (zy > 0.f) { part_dw += zy * gn.y; dpt.y += w * gn.y; }
      if (zz > 0.f) { part_dw += zz * gn.z; dpt.z += w * gn.z; }
      if (zw > 0.f) { part_dw += zw * gn.w; dpt.w += w * gn.w; }
    }
    const float dw = head_sum<LPN>(part_dw, lph) - go;
    if (on) {
      const float da = dw * w;
      dat += da;
      if (head_lead) dA[(int64_t)(e0 + i) * H + h] = da;
    }
  }
  if (!mine) return;
  if (CHUNK) {
    float* p = part + (seg * H + h) * (CH + 1) + (4 * j - h * oh);
    p[0] = dpt.x; p[1] = dpt.y; p[2] = dpt.z; p[3] = dpt.w;
    if (head_lead) part[(seg * H + h) * (CH + 1) + CH] = dat;
    return;
  }
  float* dpn = dP + n * ldp;
  *reinterpret_cast<float4*>(dpn + C + 4 * j) = dpt;
  if (head_lead) dpn[2 * C + H + h] = dat;
}

template <int LPN, bool CHUNK>
__global__ void __launch_bounds__(128) k_gat_bwd_src_v(int64_t n_seg, int H, int oh, const int32_t* __restrict__ sptr,
                                                       const int32_t* __restrict__ s_tgt, const int32_t* __restrict__ s_pos,
                                                       const float* __restrict__ P, int64_t ldp,
                                                       const unsigned long long* __restrict__ amax_key,
                                                       const float* __restrict__ den, const float* __restrict__ g, int64_t ldg,
                                                       const float* __restrict__ dA, float* __restrict__ dP, GatHeavy hv,
                                                       float* __restrict__ part, int CH) {
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t seg = gid / LPN;
  const int j = static_cast<int>(gid % LPN);
  if (seg >= n_seg) return;
  int64_t s;
  int e0, e1;
  if (!gat_segment<CHUNK>(seg, sptr, hv, s, e0, e1)) return;
  const int C = H * oh;
  const int h = (4 * j) / oh;
  const float amax = key_max(amax_key[h]);
  const float* psn = P + s * ldp;
  const float as = __ldg(psn + 2 * C + h);
  const float4 ps = ldg4(psn + 4 * j);
  float4 dps = make_float4(0.f, 0.f, 0.f, 0.f);
  float das = 0.f;
  for (int e = e0; e < e1; ++e) {
    const int64_t n = __ldg(s_tgt + e);
    const float* pt = P + n * ldp;
    const float w = expf(as + __ldg(pt + 2 * C + H + h) - amax) / __ldg(den + n * H + h);
    const float4 v = ldg4(pt + C + 4 * j), gg = ldg4(g + n * ldg + 4 * j);
    if (ps.x + v.x > 0.f) dps.x += w * gg.x;
    if (ps.y + v.y > 0.f) dps.y += w * gg.y;
    if (ps.z + v.z > 0.f) dps.z += w * gg.z;
    if (ps.w + v.w > 0.f) dps.w += w * gg.w;
    das += __ldg(dA + (int64_t)__ldg(s_pos + e) * H + h);
  }
  const bool head_lead = (4 * j) % oh == 0;
  if (CHUNK) {
    float* p = part + (seg * H + h) * (CH + 1) + (4 * j - h * oh);
    p[0] = dps.x; p[1] = dps.y; p[2] = dps.z; p[3] = dps.w;
    if (head_lead) part[(seg * H + h) * (CH + 1) + CH] = das;
    return;
  }
  float* dpn = dP + s * ldp;
  *reinterpret_cast<float4*>(dpn + 4 * j) = dps;
  if (head_lead) dpn[2 * C + h] = das;
}

// adds a hub's chunk partials in chunk order: dP[n, col0 + h*oh + j] and dP[n, colA + h]
template <int CH>
__global__ void __launch_bounds__(128) k_gat_bwd_finish(int H, int oh, GatHeavy hv, const float* __restrict__ part,
                                                        float* __restrict__ dP, int64_t ldp, int col0, int colA) {
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (gid >= (int64_t)hv.n_heavy * H) return;
  const int i = static_cast<int>(gid / H), h = static_cast<int>(gid % H);
  const int64_t n = hv.nodes[i];
  float acc[CH], da = 0.f;
#pragma unroll
  for (int j = 0; j < CH; ++j) acc[j] = 0.f;
  for (int c = hv.cptr[i]; c < hv.cptr[i + 1]; ++c) {
    const float* p = part + ((int64_t)c * H + h) * (CH + 1);
#pragma unroll
    for (int j = 0; j < CH; ++j) acc[j] += p[j];
    da += p[CH];
  }
  float* dpn = dP + n * ldp;
#pragma unroll
  for (int j = 0; j < CH; ++j)
    if (j < oh) dpn[col0 + h * oh + j] = acc[j];
  dpn[colA + h] = da;
}

// d amax[h] = -sum_e dA[e, h]  goes to the arg-max edge: das[src*] and dat[tgt*]
__global__ void k_gat_bwd_max(int H, int C, const unsigned long long* __restrict__ amax_key, const float* __restrict__ dA_colsum,
                              const int32_t* __restrict__ t_src, const int32_t* __restrict__ t_tgt, float* __restrict__ dP,
                              int64_t ldp) {
  const int h = threadIdx.x;
  if (h >= H) return;
  const unsigned long long k = amax_key[h];
  if (!k) return;
  const unsigned int pos = key_pos(k);
  const float d = -dA_colsum[h];
  // distinct heads touch distinct columns; src* != tgt* rows or distinct columns (2C+h vs 2C+H+h): no race
  dP[(int64_t)t_src[pos] * ldp + 2 * C + h] += d;
  dP[(int64_t)t_tgt[pos] * ldp + 2 * C + H + h] += d;
}

static GatHeavy heavy_of(const gode_gat_graph_t* G, bool by_source) {
  GatHeavy hv;
  const gode_gat_heavy_t& h = by_source ? G->s_heavy : G->t_heavy;
  hv.n_heavy = h.n_heavy;
  hv.n_chunks = h.n_chunks;
  hv.nodes = h.nodes;
  hv.cptr = h.cptr;
  hv.chunk_node = h.chunk_node;
  hv.chunk_e0 = h.chunk_e0;
  return hv;
}

static unsigned grid_for(int64_t work) { return static_cast<unsigned>((work + 127) / 128); }

static int ch_of(int oh) { return oh <= 8 ? 8 : oh <= 16 ? 16 : oh <= 32 ? 32 : oh <= 64 ? 64 : 128; }

// ---- lane-group path ----
static bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// lanes per segment of the lane-group kernels, or 0 when the shape / alignment needs the scalar kernels.  GODE_GAT_VEC=0
// forces the scalar kernels (tests compare the two).
static int gat_lpn(int H, int oh, int64_t ldp, const void* P, int64_t ld2, const void* p2, int64_t ld3, const void* p3) {
  const char* e = getenv("GODE_GAT_VEC");
  if (e && atoi(e) == 0) return 0;
  const int C = H * oh;
  if (oh < 4 || (oh & (oh - 1)) || C > 128) return 0;
  const int lpn = C / 4;
  if (lpn != 4 && lpn != 8 && lpn != 16 && lpn != 32) return 0;
  if (ldp % 4 || ld2 % 4 || ld3 % 4 || !al16p(P) || !al16p(p2) || !al16p(p3)) return 0;
  return lpn;
}

template <int LPN>
static int gat_fwd_v(const gode_gat_graph_t* G, int H, int oh, const float* P, int64_t ldp, float eps, float* out, int64_t ldo,
                     float* den, unsigned long long* amax_key, int* nan_flag, float* part, cudaStream_t st) {
  const GatHeavy hv = heavy_of(G, false);
  const int CH = ch_of(oh);
  k_gat_fwd_v<LPN, false><<<grid_for(G->n_nodes * LPN), 128, 0, st>>>(G->n_nodes, H, oh, G->tptr, G->t_src, P, ldp, eps, amax_key,
                                                                      out, ldo, den, nan_flag, hv, nullptr, CH);
  GODE_LAUNCH_CHECK();
  if (hv.n_chunks > 0) {
    k_gat_fwd_v<LPN, true><<<grid_for((int64_t)hv.n_chunks * LPN), 128, 0, st>>>(hv.n_chunks, H, oh, G->tptr, G->t_src, P, ldp, eps,
                                                                                 amax_key, out, ldo, den, nan_flag, hv, part, CH);
    GODE_LAUNCH_CHECK();
  }
  return GODE_OK;
}

template <int LPN>
static int gat_bwd_v(const gode_gat_graph_t* G, int H, int oh, const float* P, int64_t ldp, const float* out, int64_t ldo,
                     const float* den, const unsigned long long* amax_key, const float* g, int64_t ldg, float* dP, float* dA,
                     float* part, cudaStream_t st, int phase /*0: target pass, 1: source pass*/) {
  const int CH = ch_of(oh);
  if (phase == 0) {
    const GatHeavy ht = heavy_of(G, false);
    k_gat_bwd_tgt_v<LPN, false><<<grid_for(G->n_nodes * LPN), 128, 0, st>>>(G->n_nodes, H, oh, G->tptr, G->t_src, P, ldp, amax_key,
                                                                            out, ldo, den, g, ldg, dP, dA, ht, nullptr, CH);
    GODE_LAUNCH_CHECK();
    if (ht.n_chunks > 0) {
      k_gat_bwd_tgt_v<LPN, true><<<grid_for((int64_t)ht.n_chunks * LPN), 128, 0, st>>>(ht.n_chunks, H, oh, G->tptr, G->t_src, P, ldp,
                                                                                       amax_key, out, ldo, den, g, ldg, dP, dA, ht,
                                                                                       part, CH);
      GODE_LAUNCH_CHECK();
    }
    return GODE_OK;
  }
  const GatHeavy hs = heavy_of(G, true);
  k_gat_bwd_src_v<LPN, false><<<grid_for(G->n_nodes * LPN), 128, 0, st>>>(G->n_nodes, H, oh, G->sptr, G->s_tgt, G->s_pos, P, ldp,
                                                                          amax_key, den, g, ldg, dA, dP, hs, nullptr, CH);
  GODE_LAUNCH_CHECK();
  if (hs.n_chunks > 0) {
    k_gat_bwd_src_v<LPN, true><<<grid_for((int64_t)hs.n_chunks * LPN), 128, 0, st>>>(hs.n_chunks, H, oh, G->sptr, G->s_tgt, G->s_pos,
                                                                                     P, ldp, amax_key, den, g, ldg, dA, dP, hs, part,
                                                                                     CH);
    GODE_LAUNCH_CHECK();
  }
  return GODE_OK;
}

#define GODE_GAT_LPN_DISPATCH(lpn, CALL)       \
  switch (lpn) {                               \
    case 4: rc = CALL(4); break;               \
    case 8: rc = CALL(8); break;               \
    case 16: rc = CALL(16); break;             \
    default: rc = CALL(32); break;             \
  }


template <int CH>
static int gat_fwd_t(const gode_gat_graph_t* G, int H, int oh, const float* P, int64_t ldp, float eps, float* out, int64_t ldo,
                     float* den, unsigned long long* amax_key, int* nan_flag, float* part, cudaStream_t st) {
  const GatHeavy hv = heavy_of(G, false);
  const int lpn = gat_lpn(H, oh, ldp, P, ldo, out, 4, P);
  if (lpn) {
    int rc;
#define GODE_CALL_(L) gat_fwd_v<L>(G, H, oh, P, ldp, eps, out, ldo, den, amax_key, nan_flag, part, st)
    GODE_GAT_LPN_DISPATCH(lpn, GODE_CALL_)
#undef GODE_CALL_
    if (rc) return rc;
  } else {
    k_gat_fwd<CH, false><<<grid_for(G->n_nodes * H), 128, 0, st>>>(G->n_nodes, H, oh, G->tptr, G->t_src, P, ldp, eps, amax_key, out,
                                                                   ldo, den, nan_flag, hv, nullptr);
    GODE_LAUNCH_CHECK();
    if (hv.n_chunks > 0) {
      k_gat_fwd<CH, true><<<grid_for((int64_t)hv.n_chunks * H), 128, 0, st>>>(hv.n_chunks, H, oh, G->tptr, G->t_src, P, ldp, eps,
                                                                              amax_key, out, ldo, den, nan_flag, hv, part);
      GODE_LAUNCH_CHECK();
    }
  }
  if (hv.n_chunks > 0) {
    k_gat_fwd_finish<CH><<<grid_for((int64_t)hv.n_heavy * H), 128, 0, st>>>(H, oh, eps, hv, part, out, ldo, den, nan_flag);
    GODE_LAUNCH_CHECK();
  }
  return GODE_OK;
}

template <int CH>
static int gat_bwd_t(const gode_gat_graph_t* G, int H, int oh, const float* P, int64_t ldp, const float* out, int64_t ldo,
                     const float* den, const unsigned long long* amax_key, const float* g, int64_t ldg, float* dP, float* dA,
                     float* part, cudaStream_t st) {
  const int C = H * oh;
  const GatHeavy ht = heavy_of(G, false), hs = heavy_of(G, true);
  const int lpn = (gat_lpn(H, oh, ldp, P, ldo, out, ldg, g) && al16p(dP)) ? C / 4 : 0;
  int rc;
  if (lpn) {
#define GODE_CALL_(L) gat_bwd_v<L>(G, H, oh, P, ldp, out, ldo, den, amax_key, g, ldg, dP, dA, part, st, 0)
    GODE_GAT_LPN_DISPATCH(lpn, GODE_CALL_)
#undef GODE_CALL_
    if (rc) return rc;
  } else {
    k_gat_bwd_tgt<CH, false><<<grid_for(G->n_nodes * H), 128, 0, st>>>(G->n_nodes, H, oh, G->tptr, G->t_src, P, ldp, amax_key, out,
                                                                       ldo, den, g, ldg, dP, dA, ht, nullptr);
    GODE_LAUNCH_CHECK();
    if (ht.n_chunks > 0) {
      k_gat_bwd_tgt<CH, true><<<grid_for((int64_t)ht.n_chunks * H), 128, 0, st>>>(ht.n_chunks, H, oh, G->tptr, G->t_src, P, ldp,
                                                                                  amax_key, out, ldo, den, g, ldg, dP, dA, ht, part);
      GODE_LAUNCH_CHECK();
    }
  }
  if (ht.n_chunks > 0) {
    k_gat_bwd_finish<CH><<<grid_for((int64_t)ht.n_heavy * H), 128, 0, st>>>(H, oh, ht, part, dP, ldp, C, 2 * C + H);
    GODE_LAUNCH_CHECK();
  }
  if (lpn) {
#define GODE_CALL_(L) gat_bwd_v<L>(G, H, oh, P, ldp, out, ldo, den, amax_key, g, ldg, dP, dA, part, st, 1)
    GODE_GAT_LPN_DISPATCH(lpn, GODE_CALL_)
#undef GODE_CALL_
    if (rc) return rc;
  } else {
    k_gat_bwd_src<CH, false><<<grid_for(G->n_nodes * H), 128, 0, st>>>(G->n_nodes, H, oh, G->sptr, G->s_tgt, G->s_pos, P, ldp,
                                                                       amax_key, den, g, ldg, dA, dP, hs, nullptr);
    GODE_LAUNCH_CHECK();
    if (hs.n_chunks > 0) {
      k_gat_bwd_src<CH, true><<<grid_for((int64_t)hs.n_chunks * H), 128, 0, st>>>(hs.n_chunks, H, oh, G->sptr, G->s_tgt, G->s_pos, P,
                                                                                  ldp, amax_key, den, g, ldg, dA, dP, hs, part);
      GODE_LAUNCH_CHECK();
    }
  }
  if (hs.n_chunks > 0) {
    k_gat_bwd_finish<CH><<<grid_for((int64_t)hs.n_heavy * H), 128, 0, st>>>(H, oh, hs, part, dP, ldp, 0, 2 * C);
    GODE_LAUNCH_CHECK();
  }
  return GODE_OK;
}

static size_t part_bytes(const gode_gat_graph_t* G, int H, int oh) {
  const int64_t nc = G->t_heavy.n_chunks > G->s_heavy.n_chunks ? G->t_heavy.n_chunks : G->s_heavy.n_chunks;
  return align_up(sizeof(float) * static_cast<size_t>(nc > 0 ? nc : 1) * H * (ch_of(oh) + 1), 256);
}

}  // namespace gode

using namespace gode;

static int gat_check(const gode_gat_graph_t* G, int H, int oh, int64_t ldp) {
  GODE_REQUIRE(G && G->n_nodes >= 0 && G->n_edges >= 0, "gat: bad graph");
  GODE_REQUIRE(H >= 1 && H <= 64 && oh >= 1 && oh <= 128, "gat: heads must be in [1,64] and channels per head in [1,128]");
  GODE_REQUIRE(ldp >= 2LL * H * oh + 2 * H, "gat: ldp too small for [Ps | Pt | as | at]");
  GODE_REQUIRE(G->n_nodes == 0 || (G->tptr && G->sptr), "gat: null segment pointers");
  GODE_REQUIRE(G->n_edges == 0 || (G->t_src && G->t_tgt && G->s_tgt && G->s_pos), "gat: null edge arrays");
  for (const gode_gat_heavy_t* h : {&G->t_heavy, &G->s_heavy}) {
    GODE_REQUIRE(h->n_heavy >= 0 && h->n_chunks >= 0, "gat: bad hub table");
    GODE_REQUIRE(h->n_chunks == 0 || (h->nodes && h->cptr && h->chunk_node && h->chunk_e0), "gat: null hub table");
  }
  return GODE_OK;
}

extern "C" size_t gode_gat_fwd_workspace_bytes(const gode_gat_graph_t* G, int32_t heads, int32_t oh) {
  if (!G || heads < 1 || oh < 1) return 0;
  return part_bytes(G, heads, oh);
}

extern "C" size_t gode_gat_bwd_workspace_bytes(const gode_gat_graph_t* G, int32_t heads, int32_t oh) {
  if (!G || heads < 1 || oh < 1) return 0;
  const int64_t n_edges = G->n_edges;
  return align_up(sizeof(float) * static_cast<size_t>(n_edges > 0 ? n_edges : 1) * heads, 256) + align_up(sizeof(float) * 64, 256) +
         align_up(gode_colreduce_workspace_bytes(heads), 256) + part_bytes(G, heads, oh);
}

extern "C" int gode_gat_fwd(const gode_gat_graph_t* G, int32_t heads, int32_t oh, const float* P, int64_t ldp, float eps,
                            float* out, int64_t ldo, float* den, unsigned long long* amax_key, int32_t* nan_flag,
                            void* ws, size_t ws_bytes, void* stream) {
  int rc = gat_check(G, heads, oh, ldp);
  if (rc) return rc;
  GODE_REQUIRE(ldo >= heads * oh && amax_key && nan_flag && (G->n_nodes == 0 || (P && out && den)), "gat_fwd: bad argument");
  if (G->t_heavy.n_chunks > 0 && (!ws || ws_bytes < part_bytes(G, heads, oh))) {
    set_error("gat_fwd: workspace too small");
    return GODE_EWORKSPACE;
  }
  float* part = static_cast<float*>(ws);
  cudaStream_t st = as_stream(stream);
  GODE_CHECK_CUDA(cudaMemsetAsync(amax_key, 0, sizeof(unsigned long long) * heads, st));
  if (G->n_nodes == 0) return GODE_OK;
  if (G->n_edges > 0) {
    int64_t blocks = (G->n_edges + 255) / 256;
    const int64_t cap = 4LL * sm_count();
    if (blocks > cap) blocks = cap;
    k_gat_amax<<<static_cast<unsigned>(blocks), 256, 0, st>>>(G->n_edges, heads, G->t_src, G->t_tgt, P, ldp, heads * oh, amax_key,
                                                              nan_flag);
    GODE_LAUNCH_CHECK();
  }
  if (oh <= 8) return gat_fwd_t<8>(G, heads, oh, P, ldp, eps, out, ldo, den, amax_key, nan_flag, part, st);
  if (oh <= 16) return gat_fwd_t<16>(G, heads, oh, P, ldp, eps, out, ldo, den, amax_key, nan_flag, part, st);
  if (oh <= 32) return gat_fwd_t<32>(G, heads, oh, P, ldp, eps, out, ldo, den, amax_key, nan_flag, part, st);
  if (oh <= 64) return gat_fwd_t<64>(G, heads, oh, P, ldp, eps, out, ldo, den, amax_key, nan_flag, part, st);
  return gat_fwd_t<128>(G, heads, oh, P, ldp, eps, out, ldo, den, amax_key, nan_flag, part, st);
}

extern "C" int gode_gat_bwd(const gode_gat_graph_t* G, int32_t heads, int32_t oh, const float* P, int64_t ldp, const float* out,
                            int64_t ldo, const float* den, const unsigned long long* amax_key, const float* gout, int64_t ldg,
                            float* dP, void* ws, size_t ws_bytes, void* stream) {
  int rc = gat_check(G, heads, oh, ldp);
  if (rc) return rc;
  GODE_REQUIRE(amax_key && (G->n_nodes == 0 || (P && out && den && gout && dP)), "gat_bwd: null pointer");
  if (!ws || ws_bytes < gode_gat_bwd_workspace_bytes(G, heads, oh)) {
    set_error("gat_bwd: workspace too small");
    return GODE_EWORKSPACE;
  }
  if (G->n_nodes == 0) return GODE_OK;
  cudaStream_t st = as_stream(stream);
  Arena ar(ws, ws_bytes);
  float* dA = ar.take<float>(static_cast<size_t>(G->n_edges > 0 ? G->n_edges : 1) * heads);
  float* dsum = ar.take<float>(64);
  const size_t red_bytes = gode_colreduce_workspace_bytes(heads);
  float* red = reinterpret_cast<float*>(ar.take<char>(red_bytes));
  float* part = reinterpret_cast<float*>(ar.take<char>(part_bytes(G, heads, oh)));
  if (oh <= 8) rc = gat_bwd_t<8>(G, heads, oh, P, ldp, out, ldo, den, amax_key, gout, ldg, dP, dA, part, st);
  else if (oh <= 16) rc = gat_bwd_t<16>(G, heads, oh, P, ldp, out, ldo, den, amax_key, gout, ldg, dP, dA, part, st);
  else if (oh <= 32) rc = gat_bwd_t<32>(G, heads, oh, P, ldp, out, ldo, den, amax_key, gout, ldg, dP, dA, part, st);
  else if (oh <= 64) rc = gat_bwd_t<64>(G, heads, oh, P, ldp, out, ldo, den, amax_key, gout, ldg, dP, dA, part, st);
  else rc = gat_bwd_t<128>(G, heads, oh, P, ldp, out, ldo, den, amax_key, gout, ldg, dP, dA, part, st);
  if (rc) return rc;
  if (G->n_edges > 0) {
    if ((rc = colsum(G->n_edges, heads, dA, heads, dsum, red, red_bytes, st))) return rc;
    k_gat_bwd_max<<<1, 64, 0, st>>>(heads, heads * oh, amax_key, dsum, G->t_src, G->t_tgt, dP, ldp);
    GODE_LAUNCH_CHECK();
  }
  return GODE_OK;
}
