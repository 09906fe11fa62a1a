// libgode: CSR SpMM  Y = A * X  with a fused row epilogue (bias, ReLU, residual, Runge-Kutta stage
// combination, adjoint mask).  HBM-bound gather kernel.
//
// Reference call sites: torch.spmm(adj, support) GCN/layers.py:33,71 (+bias :35,73; F.relu GCN/models.py:178).
//
// Mapping (d = LPR*VPL*4 floats per row):
//   * one sub-warp of LPR lanes per row, each lane owns VPL float4 (128-bit) channel vectors, so a
//     gathered neighbour row is read with fully coalesced 16-byte loads (d=128: one 512 B row per warp);
//   * the row's (col,val) pairs are loaded LPR at a time, coalesced and with a streaming hint, and
//     broadcast with shuffles; neighbour-row loads are issued four at a time before use to keep >= 4
//     128-bit requests per lane in flight;
//   * CSR arrays, epilogue operands and outputs use ld/st.global.cs (evict-first) so that L2 keeps the
//     gather operand X, the only tensor with reuse;
//   * rows longer than GODE_HEAVY_ROW are skipped here and handled by k_spmm_heavy (one CTA per row).
#include "internal.cuh"
#include <string.h>

namespace gode {

struct RowCtx {
  int64_t row;
};

template <int VPL>
__device__ __forceinline__ void epilogue(const gode_spmm_epilogue_t& ep, int64_t row, int col0 /*first float of lane*/,
                                         float4 (&acc)[VPL], float* __restrict__ Y, int64_t ldy) {
#pragma unroll
  for (int u = 0; u < VPL; ++u) {
    const int c = col0 + u * 4;
    float4 v = acc[u];
    if (ep.bias) {
      float4 b = ld_ro4(ep.bias + c);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (ep.relu) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    }
    const int64_t o = row * ldy + c;
    if (Y) {
      float4 w = v;
      if (ep.residual) {
        float4 r = ld_stream4(ep.residual + o);
        w.x += r.x; w.y += r.y; w.z += r.z; w.w += r.w;
      }
      st_stream4(Y + o, w);
    }
    if (ep.ynext) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < GODE_MAX_STAGES; ++j) {
        if (j < ep.n_prev) {
          float4 k = ld_stream4(ep.kprev[j] + o);
          const float cj = ep.coef[j];
          t.x += cj * k.x; t.y += cj * k.y; t.z += cj * k.z; t.w += cj * k.w;
        }
      }
      t.x += ep.coef_self * v.x; t.y += ep.coef_self * v.y; t.z += ep.coef_self * v.z; t.w += ep.coef_self * v.w;
      float4 y = ld_stream4(ep.y0 + o);
      y.x += t.x; y.y += t.y; y.z += t.z; y.w += t.w;
      st_stream4(ep.ynext + o, y);
    }
    if (ep.gp_out) {
      float4 a = ld_stream4(ep.mask_src + o);
      float4 g;
      g.x = v.x > 0.f ? ep.mask_scale * a.x : 0.f;
      g.y = v.y > 0.f ? ep.mask_scale * a.y : 0.f;
      g.z = v.z > 0.f ? ep.mask_scale * a.z : 0.f;
      g.w = v.w > 0.f ? ep.mask_scale * a.w : 0.f;
      st_stream4(ep.gp_out + o, g);
    }
  }
}

template <int VPL>
__device__ __forceinline__ void fma_row(float4 (&acc)[VPL], float v, const float* __restrict__ xrow) {
#pragma unroll
  for (int u = 0; u < VPL; ++u) {
    float4 x = ld_ro4(xrow + u * 4);
    acc[u].x += v * x.x; acc[u].y += v * x.y; acc[u].z += v * x.z; acc[u].w += v * x.w;
  }
}

template <int LPR, int VPL>
__global__ void __launch_bounds__(256) k_spmm_vec(int64_t n_rows, const int32_t* __restrict__ rowptr,
                                                  const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                                  const float* __restrict__ X, int64_t ldx, float* __restrict__ Y,
                                                  int64_t ldy, const gode_spmm_epilogue_t ep) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const int64_t warp = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t row = warp * RPW + sub;
  const bool valid = row < n_rows;
  int e0 = 0, e1 = 0;
  if (valid) {
    e0 = __ldg(rowptr + row);
    e1 = __ldg(rowptr + row + 1);
  }
  const bool heavy = (e1 - e0) > GODE_HEAVY_ROW;
  if (heavy) e1 = e0;
  int maxlen = e1 - e0;
#pragma unroll
  for (int o = 16; o >= LPR && o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
  const int col0 = sl * VPL * 4;
  const float* __restrict__ xl = X + col0;

  float4 acc[VPL];
#pragma unroll
  for (int u = 0; u < VPL; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int off = 0; off < maxlen; off += LPR) {
    const int e = e0 + off + sl;
    int c = 0;
    float v = 0.f;
    if (e < e1) {
      c = __ldcs(colidx + e);
      v = __ldcs(vals + e);
    }
    const int cnt = min(LPR, e1 - e0 - off);  // may be <= 0 for a finished row
#pragma unroll
    for (int j = 0; j < LPR; j += 4) {
      int cj[4];
      float vj[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int src = sub * LPR + ((j + q) % LPR);
        cj[q] = __shfl_sync(0xffffffffu, c, src);
        vj[q] = __shfl_sync(0xffffffffu, v, src);
      }
      if (j + 3 < cnt) {
        float4 x[4][VPL];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int u = 0; u < VPL; ++u) x[q][u] = ld_ro4(xl + (int64_t)cj[q] * ldx + u * 4);
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int u = 0; u < VPL; ++u) {
            acc[u].x += vj[q] * x[q][u].x; acc[u].y += vj[q] * x[q][u].y;
            acc[u].z += vj[q] * x[q][u].z; acc[u].w += vj[q] * x[q][u].w;
          }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (j + q < cnt && (LPR >= 4 || q < LPR)) fma_row<VPL>(acc, vj[q], xl + (int64_t)cj[q] * ldx);
      }
    }
  }
  if (valid && !heavy) epilogue<VPL>(ep, row, col0, acc, Y, ldy);
}

// one CTA per heavy row: NS = 8*RPW "slots" stride over the row's entries, partial sums meet in smem and are
// added in slot order (deterministic)
template <int LPR, int VPL>
__global__ void __launch_bounds__(256) k_spmm_heavy(const int32_t* __restrict__ heavy_rows, const int32_t* __restrict__ rowptr,
                                                    const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                                    const float* __restrict__ X, int64_t ldx, float* __restrict__ Y,
                                                    int64_t ldy, const gode_spmm_epilogue_t ep) {
  constexpr int RPW = 32 / LPR;
  constexpr int NS = 8 * RPW;
  constexpr int D = LPR * VPL * 4;
  extern __shared__ float4 sm4[];  // [NS][D/4]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int sub = lane / LPR, sl = lane % LPR;
  const int slot = w * RPW + sub;
  const int64_t row = heavy_rows[blockIdx.x];
  const int e0 = rowptr[row], e1 = rowptr[row + 1];
  const int col0 = sl * VPL * 4;
  const float* __restrict__ xl = X + col0;
  float4 acc[VPL];
#pragma unroll
  for (int u = 0; u < VPL; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  int e = e0 + slot;
  for (; e + 3 * NS < e1; e += 4 * NS) {
    int cj[4];
    float vj[4];
    float4 x[4][VPL];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      cj[q] = __ldcs(colidx + e + q * NS);
      vj[q] = __ldcs(vals + e + q * NS);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int u = 0; u < VPL; ++u) x[q][u] = ld_ro4(xl + (int64_t)cj[q] * ldx + u * 4);
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int u = 0; u < VPL; ++u) {
        acc[u].x += vj[q] * x[q][u].x; acc[u].y += vj[q] * x[q][u].y;
        acc[u].z += vj[q] * x[q][u].z; acc[u].w += vj[q] * x[q][u].w;
      }
  }
  for (; e < e1; e += NS) fma_row<VPL>(acc, __ldcs(vals + e), xl + (int64_t)__ldcs(colidx + e) * ldx);
#pragma unroll
  for (int u = 0; u < VPL; ++u) sm4[slot * (D / 4) + sl * VPL + u] = acc[u];
  __syncthreads();
  if (slot == 0) {
#pragma unroll
    for (int u = 0; u < VPL; ++u) {
      float4 t = sm4[sl * VPL + u];
      for (int s = 1; s < NS; ++s) {
        float4 o = sm4[s * (D / 4) + sl * VPL + u];
        t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
      }
      acc[u] = t;
    }
    epilogue<VPL>(ep, row, col0, acc, Y, ldy);
  }
}

// any d (e.g. nclass = 7, QC hidden = 73): one warp per row, lanes stride over channels, scalar loads.
__global__ void __launch_bounds__(256) k_spmm_generic(int64_t n_rows, const int32_t* __restrict__ rowptr,
                                                      const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                                      const float* __restrict__ X, int64_t ldx, int d,
                                                      float* __restrict__ Y, int64_t ldy, const gode_spmm_epilogue_t ep) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int e0 = rowptr[row], e1 = rowptr[row + 1];
  for (int c0 = 0; c0 < d; c0 += 128) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int e = e0; e < e1; ++e) {
      const float v = __ldg(vals + e);
      const float* xr = X + (int64_t)__ldg(colidx + e) * ldx;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = c0 + q * 32 + lane;
        if (c < d) acc[q] += v * __ldg(xr + c);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = c0 + q * 32 + lane;
      if (c >= d) continue;
      float v = acc[q];
      if (ep.bias) v += ep.bias[c];
      if (ep.relu) v = fmaxf(v, 0.f);
      const int64_t o = row * ldy + c;
      if (Y) Y[o] = ep.residual ? v + ep.residual[o] : v;
      if (ep.ynext) {
        float t = 0.f;
        for (int j = 0; j < ep.n_prev; ++j) t += ep.coef[j] * ep.kprev[j][o];
        t += ep.coef_self * v;
        ep.ynext[o] = ep.y0[o] + t;
      }
      if (ep.gp_out) ep.gp_out[o] = v > 0.f ? ep.mask_scale * ep.mask_src[o] : 0.f;
    }
  }
}

template <int LPR, int VPL>
static int launch_vec(int64_t n_rows, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                      const int32_t* heavy_rows, int32_t n_heavy, const float* X, int64_t ldx, float* Y, int64_t ldy,
                      const gode_spmm_epilogue_t& ep, cudaStream_t st) {
  constexpr int RPW = 32 / LPR;
  constexpr int RPB = 8 * RPW;
  if (n_heavy > 0) {
    size_t smem = sizeof(float4) * RPB * LPR * VPL;
    k_spmm_heavy<LPR, VPL><<<n_heavy, 256, smem, st>>>(heavy_rows, rowptr, colidx, vals, X, ldx, Y, ldy, ep);
    GODE_LAUNCH_CHECK();
  }
  if (n_rows > 0) {
    unsigned grid = static_cast<unsigned>((n_rows + RPB - 1) / RPB);
    k_spmm_vec<LPR, VPL><<<grid, 256, 0, st>>>(n_rows, rowptr, colidx, vals, X, ldx, Y, ldy, ep);
    GODE_LAUNCH_CHECK();
  }
  return GODE_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int spmm_dispatch(int64_t n_rows, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                  const int32_t* heavy_rows, int32_t n_heavy, const float* X, int64_t ldx, int32_t d, float* Y,
                  int64_t ldy, const gode_spmm_epilogue_t& ep, cudaStream_t st) {
  bool vec_ok = (d % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(X) && aligned16(Y) && aligned16(ep.bias) &&
                aligned16(ep.residual) && aligned16(ep.y0) && aligned16(ep.ynext) && aligned16(ep.mask_src) &&
                aligned16(ep.gp_out);
  for (int j = 0; j < ep.n_prev; ++j) vec_ok = vec_ok && aligned16(ep.kprev[j]);
  if (vec_ok) {
    switch (d) {
      case 8: return launch_vec<2, 1>(n_rows, rowptr, colidx, vals, heavy_rows, n_heavy, X, ldx, Y, ldy, ep, st);
      case 16: return launch_vec<4, 1>(n_rows, rowptr, colidx, vals, heavy_rows, n_heavy, X, ldx, Y, ldy, ep, st);
      case 32: return launch_vec<8, 1>(n_rows, rowptr, colidx, vals, heavy_rows, n_heavy, X, ldx, Y, ldy, ep, st);
      case 64: return launch_vec<16, 1>(n_rows, rowptr, colidx, vals, heavy_rows, n_heavy, X, ldx, Y, ldy, ep, st);
      case 128: return launch_vec<32, 1>(n_rows, rowptr, colidx, vals, heavy_rows, n_heavy, X, ldx, Y, ldy, ep, st);
      case 256: return launch_vec<32, 2>(n_rows, rowptr, colidx, vals, heavy_rows, n_heavy, X, ldx, Y, ldy, ep, st);
      default: break;
    }
  }
  if (n_rows > 0) {
    unsigned grid = static_cast<unsigned>((n_rows + 7) / 8);
    k_spmm_generic<<<grid, 256, 0, st>>>(n_rows, rowptr, colidx, vals, X, ldx, d, Y, ldy, ep);
    GODE_LAUNCH_CHECK();
  }
  return GODE_OK;
}

}  // namespace gode

extern "C" int gode_spmm_csr_f32(int64_t n_rows, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                                 const int32_t* heavy_rows, int32_t n_heavy, const float* X, int64_t ldx, int32_t d,
                                 float* Y, int64_t ldy, const gode_spmm_epilogue_t* epi, void* stream) {
  using namespace gode;
  GODE_REQUIRE(n_rows >= 0 && d > 0 && ldx >= d && (Y == nullptr || ldy >= d), "spmm: bad shape");
  GODE_REQUIRE(rowptr && X, "spmm: null pointer");
  GODE_REQUIRE(n_heavy == 0 || heavy_rows != nullptr, "spmm: heavy row list missing");
  gode_spmm_epilogue_t ep;
  if (epi) {
    ep = *epi;
  } else {
    memset(&ep, 0, sizeof(ep));
  }
  GODE_REQUIRE(ep.n_prev >= 0 && ep.n_prev <= GODE_MAX_STAGES, "spmm: n_prev out of range");
  GODE_REQUIRE(!ep.ynext || ep.y0, "spmm: ynext needs y0");
  GODE_REQUIRE(!ep.gp_out || ep.mask_src, "spmm: gp_out needs mask_src");
  GODE_REQUIRE(Y || ep.ynext || ep.gp_out, "spmm: no output requested");
  if (!Y && ldy < d) ldy = d;  // every epilogue operand shares the leading dimension ldy
  return spmm_dispatch(n_rows, rowptr, colidx, vals, heavy_rows, n_heavy, X, ldx, d, Y, ldy, ep, as_stream(stream));
}
