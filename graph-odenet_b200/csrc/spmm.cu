// libgode: CSR SpMM  Y = A * X  with a fused row epilogue (bias, ReLU, residual, Runge-Kutta stage
// combination, adjoint mask).  HBM-bound gather kernel.
//
// Reference call sites: torch.spmm(adj, support) GCN/layers.py:33,71 (+bias :35,73; F.relu GCN/models.py:178).
//
// Mapping (d = LPR*VPL*4 floats per row):
//   * one sub-warp of LPR lanes per row, each lane owns VPL float4 (128-bit) channel vectors, so a
//     gathered neighbour row is read with fully coalesced 16-byte loads (d=128: one 512 B row per warp);
//   * the row's (col,val) pairs are loaded LPR at a time, coalesced and with a streaming hint, and
//     broadcast with shuffles; neighbour-row loads are issued UNR at a time before use to keep >= UNR
//     128-bit requests per lane in flight;
//   * CSR arrays, epilogue operands and outputs use ld/st.global.cs (evict-first) so that L2 keeps the
//     gather operand X, the only tensor with reuse;
//   * rows longer than GODE_HEAVY_ROW (power-law hubs) are skipped by the main kernel; they are cut into
//     chunks of GODE_HEAVY_CHUNK entries, each gathered by its own sub-warp (k_spmm_heavy_partial), and a
//     finishing kernel adds a row's chunks in chunk order and applies the epilogue (deterministic).
#include "internal.cuh"
#include <string.h>
#include <stdlib.h>

namespace gode {

constexpr int UNR = 8;  // neighbour rows in flight per lane

template <int VPL>
__device__ __forceinline__ void epilogue(const gode_spmm_epilogue_t& ep, int64_t row, int col0 /*first float of lane*/,
                                         float4 (&acc)[VPL], float* __restrict__ Y, int64_t ldy) {
#pragma unroll
  for (int u = 0; u < VPL; ++u) {
    const int c = col0 + u * 4;
    float4 v = acc[u];
    if (ep.acc_in) {
      float4 p = ld_stream4(ep.acc_in + row * ldy + c);
      v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
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
    if (ep.ynext || ep.second.out) {
      // one pass over (y0, k_j) serves both combinations
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 t2 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < GODE_MAX_STAGES; ++j) {
        if (j < ep.n_prev) {
          float4 k = ld_stream4(ep.kprev[j] + o);
          const float cj = ep.coef[j];
          t.x += cj * k.x; t.y += cj * k.y; t.z += cj * k.z; t.w += cj * k.w;
          if (ep.second.out) {
            const float c2 = ep.second.coef[j];
            t2.x += c2 * k.x; t2.y += c2 * k.y; t2.z += c2 * k.z; t2.w += c2 * k.w;
          }
        }
      }
      const float4 y0 = ld_stream4(ep.y0 + o);
      if (ep.ynext) {
        t.x += ep.coef_self * v.x; t.y += ep.coef_self * v.y; t.z += ep.coef_self * v.z; t.w += ep.coef_self * v.w;
        float4 y = y0;
        y.x += t.x; y.y += t.y; y.z += t.z; y.w += t.w;
        st_stream4(ep.ynext + o, y);
        if (ep.push_y.ptr) {   // fused halo push of the next stage state: the peers transform these rows themselves
          const int p1 = __ldg(ep.push_y.ptr + row + 1);
          for (int e = __ldg(ep.push_y.ptr + row); e < p1; ++e) {
            const int64_t ent = __ldg(ep.push_y.ent + e);
            float* dst = ep.push_y.base[ent >> 40] + (ent & 0xFFFFFFFFFFll) * ldy + c;
            *reinterpret_cast<float4*>(dst) = y;
          }
        }
      }
      if (ep.second.out) {
        const float cs = ep.second.coef_self;
        t2.x += cs * v.x; t2.y += cs * v.y; t2.z += cs * v.z; t2.w += cs * v.w;
        float4 y = y0;
        y.x += t2.x; y.y += t2.y; y.z += t2.z; y.w += t2.w;
        st_stream4(ep.second.out + o, y);
      }
    }
    if (ep.gp_out) {
      float4 a = ld_stream4(ep.mask_src + o);
      const float ms = ep.gp_row_scale ? ep.mask_scale * __ldg(ep.gp_row_scale + row) : ep.mask_scale;
      float4 g;
      g.x = v.x > 0.f ? ms * a.x : 0.f;
      g.y = v.y > 0.f ? ms * a.y : 0.f;
      g.z = v.z > 0.f ? ms * a.z : 0.f;
      g.w = v.w > 0.f ? ms * a.w : 0.f;
      st_stream4(ep.gp_out + o, g);
      if (ep.push.ptr) {   // fused halo push: the peers that reference this row get it now, over NVLink
        const int p1 = __ldg(ep.push.ptr + row + 1);
        for (int e = __ldg(ep.push.ptr + row); e < p1; ++e) {
          const int64_t ent = __ldg(ep.push.ent + e);
          float* dst = ep.push.base[ent >> 40] + (ent & 0xFFFFFFFFFFll) * ldy + c;
          *reinterpret_cast<float4*>(dst) = g;
        }
      }
    }
  }
}

// The epilogue's row operands (y0, k_j, a, ...) are DRAM-resident streams that a warp only needs after its gather.
// Asking L2 for them before the gather starts turns their ~1 us DRAM latency at the end of the warp's life into an
// L2 hit, so warp slots recycle sooner; no registers are held (prefetch.global.L2).
__device__ __forceinline__ void prefetch_l2(const float* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int VPL>
__device__ __forceinline__ void epilogue_prefetch(const gode_spmm_epilogue_t& ep, int64_t row, int col0, int64_t ldy) {
  const int64_t o = row * ldy + col0;
#pragma unroll
  for (int u = 0; u < VPL; ++u) {
    const int64_t ou = o + u * 4;
    if (ep.acc_in) prefetch_l2(ep.acc_in + ou);
    if (ep.residual) prefetch_l2(ep.residual + ou);
    if (ep.ynext || ep.second.out) {
      prefetch_l2(ep.y0 + ou);
#pragma unroll
      for (int j = 0; j < GODE_MAX_STAGES; ++j)
        if (j < ep.n_prev) prefetch_l2(ep.kprev[j] + ou);
    }
    if (ep.gp_out) prefetch_l2(ep.mask_src + ou);
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

// acc += sum_{e in [e0,e1)} vals[e] * X[colidx[e], lane's channels]; all 32 lanes of the warp call this together
// (sub-warps own different ranges; `maxlen` is the longest range in the warp).
template <int LPR, int VPL, int UN = UNR>
__device__ __forceinline__ void gather_range(float4 (&acc)[VPL], int e0, int e1, int maxlen, int sub, int sl,
                                             const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                             const float* __restrict__ xl, int64_t ldx) {
  constexpr int U = UN < LPR ? UN : LPR;
  for (int off = 0; off < maxlen; off += LPR) {
    const int e = e0 + off + sl;
    int c = 0;
    float v = 0.f;
    if (e < e1) {
      c = __ldcs(colidx + e);
      v = __ldcs(vals + e);
    }
    const int cnt = min(LPR, e1 - e0 - off);  // may be <= 0 for a finished range
    // Entries past the end of the range are padded with the range's own first column and weight 0, so that the last,
    // partial batch is gathered with U independent loads like every other one (a sequential tail -- one dependent load
    // after the other -- was what made 4 rows in flight beat 8: 223.5 vs 235.0 ms per step at N = 10 M).  The padded
    // loads repeat a line this warp has just requested; multiplying a real neighbour's row by 0 cannot introduce a
    // NaN the exact sum would not have.
    const int cpad = __shfl_sync(0xffffffffu, c, sub * LPR);
#pragma unroll
    for (int j = 0; j < LPR; j += U) {
      int cj[U];
      float vj[U];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        cj[q] = __shfl_sync(0xffffffffu, c, sub * LPR + j + q);
        vj[q] = __shfl_sync(0xffffffffu, v, sub * LPR + j + q);
      }
      if (j < cnt) {
        float4 x[U][VPL];
#pragma unroll
        for (int q = 0; q < U; ++q) {
          const int cq = (j + q < cnt) ? cj[q] : cpad;
#pragma unroll
          for (int u = 0; u < VPL; ++u) x[q][u] = ld_ro4(xl + (int64_t)cq * ldx + u * 4);
        }
#pragma unroll
        for (int q = 0; q < U; ++q)
#pragma unroll
          for (int u = 0; u < VPL; ++u) {
            acc[u].x += vj[q] * x[q][u].x; acc[u].y += vj[q] * x[q][u].y;
            acc[u].z += vj[q] * x[q][u].z; acc[u].w += vj[q] * x[q][u].w;
          }
      }
    }
  }
}

template <int LPR>
__device__ __forceinline__ int warp_max_over_subs(int v) {
#pragma unroll
  for (int o = 16; o >= LPR; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <int LPR, int VPL, int MINB = 1, int UN = UNR>
__global__ void __launch_bounds__(256, MINB) k_spmm_vec(int64_t n_rows, const int32_t* __restrict__ rowptr,
                                                  const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                                  const float* __restrict__ X, int64_t ldx, float* __restrict__ Y,
                                                  int64_t ldy, const gode_spmm_epilogue_t ep, const int prefetch) {
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
  const int maxlen = warp_max_over_subs<LPR>(e1 - e0);
  const int col0 = sl * VPL * 4;
  if (prefetch && valid && !heavy) epilogue_prefetch<VPL>(ep, row, col0, ldy);
  float4 acc[VPL];
#pragma unroll
  for (int u = 0; u < VPL; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  gather_range<LPR, VPL, UN>(acc, e0, e1, maxlen, sub, sl, colidx, vals, X + col0, ldx);
  if (valid && !heavy) epilogue<VPL>(ep, row, col0, acc, Y, ldy);
}


// ------------------------------------------------------------------------------------------------
// k_spmm_sb: the gather for one WARP per row (d = 128 / 256) with the column indices broadcast through SHARED MEMORY.
//
// Round 2 microbenchmark (tools/l2_gather_bench.cu, profiles/r02_l2_gather.jsonl): with no CSR at all this B200 delivers
// random 512 B rows from L2 to the SMs at 16.9 - 17.9 TB/s (= 62 B/clk/SM, the L2 -> L1 fill port), while k_spmm_vec
// reaches 11.2 TB/s with its L1 LSU data pipe 85 % busy.  The difference is the index broadcast: k_spmm_vec holds a
// row's (col, val) pairs one per lane and broadcasts them with two SHFLs per neighbour row, and a SHFL occupies the same
// LSU pipe as the 4 wavefronts of the 512 B row load it feeds.  Here the 32 pairs of a batch are stored to the warp's
// 256-byte slot of shared memory once (1 + 1 wavefronts) and read back four at a time with broadcast LDS.128 (one
// wavefront per 4 neighbours): 10 LSU wavefronts of index traffic per 32 neighbours instead of 64.
//   ROWVAL: the matrix is row-constant (gode_csr_t.row_vals; the reference's D^-1 (A + I) is) -> the values stream is not
//   read at all and the row sum is scaled once.
// ------------------------------------------------------------------------------------------------
template <int VPL, int UN, bool ROWVAL>
__device__ __forceinline__ void gather_range_sb(float4 (&acc)[VPL], int e0, int e1, int lane, int* __restrict__ sidx,
                                                float* __restrict__ sval, const int32_t* __restrict__ colidx,
                                                const float* __restrict__ vals, const float* __restrict__ xl, int64_t ldx) {
  for (int off = e0; off < e1; off += 32) {
    const int e = off + lane;
    int c = 0;
    float v = 0.f;
    if (e < e1) {
      c = __ldcs(colidx + e);
      if (!ROWVAL) v = __ldcs(vals + e);
    }
    __syncwarp();                       // every lane has finished reading the previous batch
    sidx[lane] = c;
    if (!ROWVAL) sval[lane] = v;
    __syncwarp();
    const int cnt = min(32, e1 - off);
#pragma unroll 2
    for (int j = 0; j < cnt; j += UN) {
      int cj[UN];
      float vj[UN];
#pragma unroll
      for (int q = 0; q < UN; q += 4) {
        const int4 t = *reinterpret_cast<const int4*>(sidx + j + q);
        cj[q] = t.x; cj[q + 1] = t.y; cj[q + 2] = t.z; cj[q + 3] = t.w;
        if (!ROWVAL) {
          const float4 w = *reinterpret_cast<const float4*>(sval + j + q);
          vj[q] = w.x; vj[q + 1] = w.y; vj[q + 2] = w.z; vj[q + 3] = w.w;
        }
      }
      float4 x[UN][VPL];
#pragma unroll
      for (int q = 0; q < UN; ++q) {
        const bool on = j + q < cnt;    // predicated-off loads are not issued; the batch still goes out back to back
#pragma unroll
        for (int u = 0; u < VPL; ++u) {
          x[q][u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (on) x[q][u] = ld_ro4(xl + (int64_t)cj[q] * ldx + u * 128);
        }
      }
#pragma unroll
      for (int q = 0; q < UN; ++q)
#pragma unroll
        for (int u = 0; u < VPL; ++u) {
          if (ROWVAL) {
            acc[u].x += x[q][u].x; acc[u].y += x[q][u].y; acc[u].z += x[q][u].z; acc[u].w += x[q][u].w;
          } else {   // entries past the range carry weight 0 (and x = 0)
            acc[u].x += vj[q] * x[q][u].x; acc[u].y += vj[q] * x[q][u].y;
            acc[u].z += vj[q] * x[q][u].z; acc[u].w += vj[q] * x[q][u].w;
          }
        }
    }
  }
}

// lane l owns channels [4l, 4l+4) of every 128-channel block (VPL blocks): a neighbour row is VPL coalesced 512 B requests
template <int VPL>
__device__ __forceinline__ void epilogue_sb(const gode_spmm_epilogue_t& ep, int64_t row, int lane, float4 (&acc)[VPL],
                                            float* __restrict__ Y, int64_t ldy) {
#pragma unroll
  for (int u = 0; u < VPL; ++u) {
    float4 one[1] = {acc[u]};
    epilogue<1>(ep, row, lane * 4 + u * 128, one, Y, ldy);
  }
}

template <int VPL, int MINB, int UN, bool ROWVAL>
__global__ void __launch_bounds__(256, MINB) k_spmm_sb(int64_t n_rows, const int32_t* __restrict__ rowptr,
                                                       const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                                       const float* __restrict__ row_vals, const float* __restrict__ X,
                                                       int64_t ldx, float* __restrict__ Y, int64_t ldy,
                                                       const gode_spmm_epilogue_t ep, const int prefetch) {
  __shared__ __align__(16) int s_idx[8][32];
  __shared__ __align__(16) float s_val[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = blockIdx.x * (int64_t)8 + w;
  if (row >= n_rows) return;            // whole warp
  const int e0 = __ldg(rowptr + row);
  int e1 = __ldg(rowptr + row + 1);
  const bool heavy = (e1 - e0) > GODE_HEAVY_ROW;
  if (heavy) return;
  if (prefetch) {
#pragma unroll
    for (int u = 0; u < VPL; ++u) epilogue_prefetch<1>(ep, row, lane * 4 + u * 128, ldy);
  }
  float4 acc[VPL];
#pragma unroll
  for (int u = 0; u < VPL; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  gather_range_sb<VPL, UN, ROWVAL>(acc, e0, e1, lane, s_idx[w], s_val[w], colidx, vals, X + lane * 4, ldx);
  if (ROWVAL) {
    const float rv = __ldg(row_vals + row);
#pragma unroll
    for (int u = 0; u < VPL; ++u) { acc[u].x *= rv; acc[u].y *= rv; acc[u].z *= rv; acc[u].w *= rv; }
  }
  epilogue_sb<VPL>(ep, row, lane, acc, Y, ldy);
}

// chunks of the heavy rows with the same gather -> partial[chunk][d] (values always applied: partial sums of a row-constant
// matrix would need the row value in the finishing kernel)
template <int VPL, int UN>
__global__ void __launch_bounds__(256) k_spmm_sb_heavy(int n_heavy, int n_chunks, const int32_t* __restrict__ heavy_rows,
                                                       const int32_t* __restrict__ chunk_ptr, const int32_t* __restrict__ rowptr,
                                                       const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                                       const float* __restrict__ X, int64_t ldx, float* __restrict__ partial) {
  __shared__ __align__(16) int s_idx[8][32];
  __shared__ __align__(16) float s_val[8][32];
  constexpr int D = VPL * 128;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int chunk = blockIdx.x * 8 + w;
  if (chunk >= n_chunks) return;
  int lo = 0, hi = n_heavy - 1;  // last heavy row whose first chunk is <= chunk
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(chunk_ptr + mid) <= chunk) lo = mid; else hi = mid - 1;
  }
  const int row = __ldg(heavy_rows + lo);
  const int r0 = __ldg(rowptr + row), r1 = __ldg(rowptr + row + 1);
  const int e0 = r0 + (chunk - __ldg(chunk_ptr + lo)) * GODE_HEAVY_CHUNK;
  const int e1 = min(r1, e0 + GODE_HEAVY_CHUNK);
  float4 acc[VPL];
#pragma unroll
  for (int u = 0; u < VPL; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  gather_range_sb<VPL, UN, false>(acc, e0, e1, lane, s_idx[w], s_val[w], colidx, vals, X + lane * 4, ldx);
#pragma unroll
  for (int u = 0; u < VPL; ++u) *reinterpret_cast<float4*>(partial + (int64_t)chunk * D + lane * 4 + u * 128) = acc[u];
}


// ------------------------------------------------------------------------------------------------
// k_spmm_ts: row-TILE staged gather (d = 128).  Round 2 finding: with its index broadcast made cheaper (k_spmm_sb above) the
// gather got SLOWER, so the L1 LSU pipe is not what separates k_spmm_vec (11 TB/s of neighbour rows) from the no-CSR
// microbenchmark (17 TB/s): it is the dependent chain rowptr -> (col, val) -> neighbour rows that every warp pays for every
// row -- two DRAM latencies before the first useful byte, for a median row of far fewer than 20 entries on a power-law
// graph.  Here a CTA owns TS_ROWS consecutive rows: all 256 threads first copy the tile's rowptr slice and its (col, val)
// entries into shared memory with coalesced loads (ONE exposed latency per tile, hidden by the other resident CTAs), then
// the 8 warps pull rows from a shared counter (long and short rows balance out) and gather with indices read from shared
// memory by broadcast LDS.  CTAs are still dispatched in row order, so the L2 window of a locality-ordered graph is kept
// (the persistent-warp variant lost it).  Rows that do not fit the staging capacity read their indices from global memory.
// ------------------------------------------------------------------------------------------------
// TS_ROWS rows per tile, staging capacity 32 entries per row (avg 19.6 on the bench graph).  pf_dist > 0: a warp that has
// run out of rows asks L2 for the neighbour rows of the tile pf_dist tiles ahead (prefetch.global.L2 on its share of that
// tile's column indices), so that the ~30 % of gathers that miss L2 today find their line already on its way.
template <int TS_ROWS, int MINB, int UN, bool ROWVAL>
__global__ void __launch_bounds__(256, MINB) k_spmm_ts(int64_t n_rows, const int32_t* __restrict__ rowptr,
                                                       const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                                       const float* __restrict__ row_vals, const float* __restrict__ X,
                                                       int64_t ldx, float* __restrict__ Y, int64_t ldy,
                                                       const gode_spmm_epilogue_t ep, const int prefetch, const int pf_dist) {
  constexpr int TS_CAP = TS_ROWS * 32;
  __shared__ int s_ptr[TS_ROWS + 1];
  __shared__ int s_idx[TS_CAP];
  __shared__ float s_val[ROWVAL ? 1 : TS_CAP];
  __shared__ int s_next;
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t row0 = blockIdx.x * (int64_t)TS_ROWS;
  const int nr = static_cast<int>(min((int64_t)TS_ROWS, n_rows - row0));
  if (tid <= nr) s_ptr[tid] = __ldg(rowptr + row0 + tid);
  if (tid == 0) s_next = 0;
  __syncthreads();
  const int t0 = s_ptr[0];
  const int nst = min(s_ptr[nr] - t0, TS_CAP);
  for (int i = tid; i < nst; i += 256) {
    s_idx[i] = __ldcs(colidx + t0 + i);
    if (!ROWVAL) s_val[i] = __ldcs(vals + t0 + i);
  }
  __syncthreads();
  const float* __restrict__ xl = X + lane * 4;
  for (;;) {
    int r = 0;
    if (lane == 0) r = atomicAdd(&s_next, 1);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r >= nr) break;
    const int e0 = s_ptr[r], e1 = s_ptr[r + 1];
    if (e1 - e0 > GODE_HEAVY_ROW) continue;          // hub rows: k_spmm_heavy_partial / finish
    const int64_t row = row0 + r;
    if (prefetch) epilogue_prefetch<1>(ep, row, lane * 4, ldy);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool staged = e1 - t0 <= nst;              // this row's entries are all in shared memory
#pragma unroll 2
    for (int j = e0; j < e1; j += UN) {
      int cj[UN];
      float vj[UN];
#pragma unroll
      for (int q = 0; q < UN; ++q) {
        const bool on = j + q < e1;
        cj[q] = 0;
        vj[q] = 0.f;
        if (on) {
          if (staged) {
            cj[q] = s_idx[j + q - t0];
            if (!ROWVAL) vj[q] = s_val[j + q - t0];
          } else {
            cj[q] = __ldg(colidx + j + q);
            if (!ROWVAL) vj[q] = __ldg(vals + j + q);
          }
        }
      }
      float4 x[UN];
#pragma unroll
      for (int q = 0; q < UN; ++q) {
        x[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j + q < e1) x[q] = ld_ro4(xl + (int64_t)cj[q] * ldx);
      }
#pragma unroll
      for (int q = 0; q < UN; ++q) {
        if (ROWVAL) {
          acc.x += x[q].x; acc.y += x[q].y; acc.z += x[q].z; acc.w += x[q].w;
        } else {
          acc.x += vj[q] * x[q].x; acc.y += vj[q] * x[q].y; acc.z += vj[q] * x[q].z; acc.w += vj[q] * x[q].w;
        }
      }
    }
    if (ROWVAL) {
      const float rv = __ldg(row_vals + row);
      acc.x *= rv; acc.y *= rv; acc.z *= rv; acc.w *= rv;
    }
    float4 one[1] = {acc};
    epilogue<1>(ep, row, lane * 4, one, Y, ldy);
  }
  if (pf_dist > 0) {
    const int64_t frow0 = row0 + (int64_t)pf_dist * TS_ROWS;
    if (frow0 < n_rows) {
      const int64_t frow1 = min(frow0 + TS_ROWS, n_rows);
      const int f0 = __ldg(rowptr + frow0);
      const int f1 = min(__ldg(rowptr + frow1), f0 + TS_CAP);
      for (int i = f0 + tid; i < f1; i += 256) {
        const float* r = X + (int64_t)__ldcs(colidx + i) * ldx;
        prefetch_l2(r);
        prefetch_l2(r + 32);
        prefetch_l2(r + 64);
        prefetch_l2(r + 96);
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------
// k_spmm_t2: k_spmm_ts with a LEAN inner loop, for contiguous 128-float rows (ldx = 128).
//
// ncu on k_spmm_ts (profiles/r02_gather.md): issue slots 85 % busy, sm__throughput 85 %, L1 / L2 / DRAM all below 65 % -- the
// gather was INSTRUCTION-bound: 33 warp instructions per gathered neighbour row, most of them 64-bit address arithmetic on
// the run-time leading dimension (IMAD x3 + LEA + LEA.HI.X per row), per-entry predicates and staged / unstaged selects.
// Here the leading dimension is the compile-time 128, so a neighbour row's address is ONE IMAD.WIDE.U32 (base + col * 512);
// full batches of four run without predicates (one predicated tail batch per row); rows that do not fit the staging
// capacity take a separate (rare) path; and the accumulation uses Blackwell's packed fp32 instructions (FADD2 / FFMA2:
// two of the four channels a lane owns per instruction) -- the same additions in the same order, so the result is
// bit-identical to the scalar form.  About 8 instructions per neighbour row.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void add2(float2& a, const float x, const float y) {
  unsigned long long ua = *reinterpret_cast<unsigned long long*>(&a);
  const float2 b = make_float2(x, y);
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(ua) : "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  a = *reinterpret_cast<float2*>(&ua);
}
__device__ __forceinline__ void fma2(float2& a, const float v, const float x, const float y) {
  unsigned long long ua = *reinterpret_cast<unsigned long long*>(&a);
  const float2 b = make_float2(x, y), vv = make_float2(v, v);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(ua) : "l"(*reinterpret_cast<const unsigned long long*>(&vv)),
      "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  a = *reinterpret_cast<float2*>(&ua);
}

template <bool ROWVAL>
__device__ __forceinline__ void acc_row(float2& a0, float2& a1, const float v, const float4& x) {
  if (ROWVAL) {
    add2(a0, x.x, x.y);
    add2(a1, x.z, x.w);
  } else {
    fma2(a0, v, x.x, x.y);
    fma2(a1, v, x.z, x.w);
  }
}

// Hub rows (> GODE_HEAVY_ROW entries) are cut into 256-entry chunks gathered by single warps (partial[chunk][128]) and summed in
// chunk order by k_spmm_heavy_finish (deterministic, no atomics).  Round 2 ncu on the SEPARATE hub kernel (k_spmm_heavy2,
// profiles/r02_gather.md): L2 hit rate 18 %, 11.7 GB of DRAM reads for 15.6 GB gathered -- swept on their own, the hubs' chunks
// find nothing of their band in L2 (the reuse distance between two hubs that share neighbours is about the size of L2), while
// the tile sweep of the light rows has exactly that band resident when it passes a hub's id.  Gathering a hub inside the CTA of
// the tile that contains it was measured slower (8.55 vs 7.08 ms per bare gather: the CTA parks its 8 warps behind two block
// barriers and holds its SM slot for the hub's whole length).  So the hub chunks become CTAs OF THE SAME GRID instead: the plan's
// work order (gode_csr_t.tile_sched) lists the 32-row tiles in id order with every group of 8 hub chunks inserted right after the
// tile that contains its hub -- the block scheduler issues CTAs in blockIdx order, so a chunk group runs while the tiles around
// its hub (and with them the hub's band) are in flight; a 100 000-entry hub is 49 consecutive CTAs spread over the SMs.
struct HubArgs {
  const int32_t* sched;       // gode_csr_t.tile_sched or nullptr (blockIdx = tile; hubs left to k_spmm_heavy2)
  const int32_t* rows;        // sorted hub row ids (gode_csr_t.heavy_rows)
  const int32_t* chunk_ptr;   // first chunk of each hub in the workspace (gode_csr_t.heavy_chunk_ptr)
  int n_heavy, n_chunks;
  float4* partial;            // [n_chunks][32] float4 workspace
};

template <int TS_ROWS, int MINB, bool ROWVAL>
__global__ void __launch_bounds__(256, MINB) k_spmm_t2(int64_t n_rows, const int32_t* __restrict__ rowptr,
                                                       const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                                       const float* __restrict__ row_vals, const float4* __restrict__ X4,
                                                       float* __restrict__ Y, const gode_spmm_epilogue_t ep, const int prefetch,
                                                       const HubArgs hub, const int64_t row_begin /*rows [row_begin, n_rows)*/) {
  constexpr int TS_CAP = TS_ROWS * 32;
  static_assert(TS_CAP >= 8 * 128, "the hub phase stages 128 entries per warp in the tile's index buffer");
  __shared__ int s_ptr[TS_ROWS + 1];
  __shared__ int s_idx[TS_CAP];
  __shared__ float s_val[ROWVAL ? 1 : TS_CAP];
  __shared__ int s_next;
  const int tid = threadIdx.x, lane = tid & 31;
  const float4* __restrict__ xb = X4 + lane;        // row c, this lane's four channels: xb[c * 32]
  int item = blockIdx.x;
  if (hub.sched) {
    item = __ldg(hub.sched + blockIdx.x);
    if (item < 0) {                                  // (uniform) a group of 8 hub chunks, one per warp
      const int w = tid >> 5;
      const int chunk = (-item - 1) * 8 + w;
      if (chunk >= hub.n_chunks) return;
      // the hub row this chunk belongs to = last one whose first chunk is <= chunk: a 32-way search (as k_spmm_heavy2)
      int lo = 0, hi = hub.n_heavy;
      while (hi - lo > 1) {
        const int step = (hi - lo + 31) >> 5;
        const int p = lo + lane * step;
        const bool le = p < hi && __ldg(hub.chunk_ptr + p) <= chunk;
        const int k = __popc(__ballot_sync(0xffffffffu, le)) - 1;
        lo += k * step;
        hi = min(lo + step, hi);
      }
      const int row = __ldg(hub.rows + lo);
      const int r1 = __ldg(rowptr + row + 1);
      const int ce0 = __ldg(rowptr + row) + (chunk - __ldg(hub.chunk_ptr + lo)) * GODE_HEAVY_CHUNK;
      const int cnt = min(r1 - ce0, GODE_HEAVY_CHUNK);
      int* __restrict__ my_idx = s_idx + w * 128;    // 128 entries at a time through this warp's slice of the index buffer
      float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
      for (int half = 0; half < cnt; half += 128) {
        const int hc = min(128, cnt - half);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = q * 32 + lane;
          if (i < hc) my_idx[i] = __ldcs(colidx + ce0 + half + i);
        }
        __syncwarp();
        // ROWVAL: plain row sums, scaled once below (as the light rows are) -- no values stream, and the four row loads of a
        // batch stay in flight together within the 32-register budget; else values straight from global memory (broadcast
        // loads of one line)
        const float* __restrict__ vp = vals + ce0 + half;
        int i = 0;
#pragma unroll 1
        for (; i + 4 <= hc; i += 4) {
          const unsigned c0 = my_idx[i], c1 = my_idx[i + 1], c2 = my_idx[i + 2], c3 = my_idx[i + 3];
          const float4 x0 = __ldg(xb + (size_t)c0 * 32), x1 = __ldg(xb + (size_t)c1 * 32);
          const float4 x2 = __ldg(xb + (size_t)c2 * 32), x3 = __ldg(xb + (size_t)c3 * 32);
          float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
          if (!ROWVAL) { v0 = __ldg(vp + i); v1 = __ldg(vp + i + 1); v2 = __ldg(vp + i + 2); v3 = __ldg(vp + i + 3); }
          acc_row<ROWVAL>(a0, a1, v0, x0);
          acc_row<ROWVAL>(a0, a1, v1, x1);
          acc_row<ROWVAL>(a0, a1, v2, x2);
          acc_row<ROWVAL>(a0, a1, v3, x3);
        }
        for (; i < hc; ++i) acc_row<ROWVAL>(a0, a1, ROWVAL ? 0.f : __ldg(vp + i), __ldg(xb + (size_t)(unsigned)my_idx[i] * 32));
      }
      if (ROWVAL) {
        const float rv = __ldg(row_vals + row);
        a0.x *= rv; a0.y *= rv; a1.x *= rv; a1.y *= rv;
      }
      hub.partial[(size_t)chunk * 32 + lane] = make_float4(a0.x, a0.y, a1.x, a1.y);
      return;
    }
  }
  const int64_t row0 = row_begin + item * (int64_t)TS_ROWS;
  const int nr = static_cast<int>(min((int64_t)TS_ROWS, n_rows - row0));
  if (tid <= nr) s_ptr[tid] = __ldg(rowptr + row0 + tid);
  if (tid == 0) s_next = 0;
  __syncthreads();
  const int t0 = s_ptr[0];
  const int nst = min(s_ptr[nr] - t0, TS_CAP);
  for (int i = tid; i < nst; i += 256) {
    s_idx[i] = __ldcs(colidx + t0 + i);
    if (!ROWVAL) s_val[i] = __ldcs(vals + t0 + i);
  }
  __syncthreads();
  for (;;) {
    int r = 0;
    if (lane == 0) r = atomicAdd(&s_next, 1);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r >= nr) break;
    const int e0 = s_ptr[r], e1 = s_ptr[r + 1];
    if (e1 - e0 > GODE_HEAVY_ROW) continue;          // hub row: its chunks are other CTAs' (or k_spmm_heavy2's) work, then k_spmm_heavy_finish
    const int64_t row = row0 + r;
    if (prefetch) epilogue_prefetch<1>(ep, row, lane * 4, 128);
    float4 one[1];
    if (e1 - t0 <= nst) {                            // the row's entries are staged (all but the tail of an over-full tile)
      float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
      int j = e0 - t0;
      const int jend = e1 - t0;
#pragma unroll 1
      for (; j + 4 <= jend; j += 4) {
        const unsigned c0 = s_idx[j], c1 = s_idx[j + 1], c2 = s_idx[j + 2], c3 = s_idx[j + 3];
        const float4 x0 = __ldg(xb + (size_t)c0 * 32), x1 = __ldg(xb + (size_t)c1 * 32);
        const float4 x2 = __ldg(xb + (size_t)c2 * 32), x3 = __ldg(xb + (size_t)c3 * 32);
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
        if (!ROWVAL) { v0 = s_val[j]; v1 = s_val[j + 1]; v2 = s_val[j + 2]; v3 = s_val[j + 3]; }
        acc_row<ROWVAL>(a0, a1, v0, x0);
        acc_row<ROWVAL>(a0, a1, v1, x1);
        acc_row<ROWVAL>(a0, a1, v2, x2);
        acc_row<ROWVAL>(a0, a1, v3, x3);
      }
      const int rem = jend - j;                      // 0..3 entries left: one predicated batch
      if (rem > 0) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 x0 = z, x1 = z, x2 = z;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
        x0 = __ldg(xb + (size_t)(unsigned)s_idx[j] * 32);
        if (rem > 1) x1 = __ldg(xb + (size_t)(unsigned)s_idx[j + 1] * 32);
        if (rem > 2) x2 = __ldg(xb + (size_t)(unsigned)s_idx[j + 2] * 32);
        if (!ROWVAL) {
          v0 = s_val[j];
          if (rem > 1) v1 = s_val[j + 1];
          if (rem > 2) v2 = s_val[j + 2];
        }
        acc_row<ROWVAL>(a0, a1, v0, x0);
        if (rem > 1) acc_row<ROWVAL>(a0, a1, v1, x1);
        if (rem > 2) acc_row<ROWVAL>(a0, a1, v2, x2);
      }
      if (ROWVAL) {
        const float rv = __ldg(row_vals + row);
        a0.x *= rv; a0.y *= rv; a1.x *= rv; a1.y *= rv;
      }
      one[0] = make_float4(a0.x, a0.y, a1.x, a1.y);
    } else {                                         // rare: indices straight from global memory (shuffle-broadcast path)
      one[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      gather_range<32, 1, 4>(one, e0, e1, e1 - e0, 0, lane, colidx, vals, reinterpret_cast<const float*>(xb), 128);
    }
    epilogue<1>(ep, row, lane * 4, one, Y, 128);
  }
}


// k_spmm_t3: k_spmm_t2 with every row's staged entries starting at a multiple of four, so that a batch's four column
// indices (and values) are ONE broadcast LDS.128 each instead of four LDS.32.  ncu on k_spmm_t2 inside the bench
// (profiles/r02_gather.md): L1 LSU data pipe 89 % busy, issue slots 66 % -- after the instruction diet the gather sits on
// the LSU wavefront rate: per 4 neighbour rows 16 wavefronts of row data + 4 (+4) of index (value) reads; aligned staging
// makes that 16 + 1 (+1).  Tiles are 32 rows (one warp computes the padded prefix of the row lengths); hub rows are not
// staged (their own kernel gathers them); a row that does not fit the staging capacity takes the shuffle-broadcast path.
template <int MINB, bool ROWVAL>
__global__ void __launch_bounds__(256, MINB) k_spmm_t3(int64_t n_rows, const int32_t* __restrict__ rowptr,
                                                       const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                                       const float* __restrict__ row_vals, const float4* __restrict__ X4,
                                                       float* __restrict__ Y, const gode_spmm_epilogue_t ep, const int prefetch) {
  constexpr int R = 32;
  constexpr int CAP = R * 32 + 128;                 // staged entries incl. padding (avg tile: 32 x 19.6 + 32 x 1.5)
  __shared__ int s_ptr[R + 1];
  __shared__ int s_pos[R];                           // first staged slot of each row (multiple of 4), -1: not staged
  __shared__ __align__(16) int s_idx[CAP];
  __shared__ __align__(16) float s_val[ROWVAL ? 4 : CAP];
  __shared__ int s_next;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t row0 = blockIdx.x * (int64_t)R;
  const int nr = static_cast<int>(min((int64_t)R, n_rows - row0));
  if (tid <= nr) s_ptr[tid] = __ldg(rowptr + row0 + tid);
  if (tid == 0) s_next = 0;
  __syncthreads();
  if (w == 0) {                                      // padded exclusive prefix of the row lengths
    int len = lane < nr ? s_ptr[lane + 1] - s_ptr[lane] : 0;
    if (len > GODE_HEAVY_ROW) len = 0;
    const int pad = (len + 3) & ~3;
    int incl = pad;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    s_pos[lane] = (incl <= CAP) ? incl - pad : -1;
  }
  __syncthreads();
  for (int r = w; r < nr; r += 8) {                  // a warp stages whole rows: coalesced reads, aligned row starts
    const int pos = s_pos[r];
    const int e0 = s_ptr[r], len = s_ptr[r + 1] - e0;
    if (pos < 0 || len > GODE_HEAVY_ROW) continue;
    for (int k = lane; k < len; k += 32) {
      s_idx[pos + k] = __ldcs(colidx + e0 + k);
      if (!ROWVAL) s_val[pos + k] = __ldcs(vals + e0 + k);
    }
  }
  __syncthreads();
  const float4* __restrict__ xb = X4 + lane;        // row c, this lane's four channels: xb[c * 32]
  for (;;) {
    int r = 0;
    if (lane == 0) r = atomicAdd(&s_next, 1);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r >= nr) break;
    const int e0 = s_ptr[r], len = s_ptr[r + 1] - e0;
    if (len > GODE_HEAVY_ROW) continue;              // hub rows: k_spmm_heavy2 / finish
    const int64_t row = row0 + r;
    if (prefetch) epilogue_prefetch<1>(ep, row, lane * 4, 128);
    float4 one[1];
    const int pos = s_pos[r];
    if (pos >= 0) {
      float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
      const int4* __restrict__ ci = reinterpret_cast<const int4*>(s_idx + pos);
      const float4* __restrict__ vi = reinterpret_cast<const float4*>(s_val + (ROWVAL ? 0 : pos));
      const int nb = len >> 2;
#pragma unroll 1
      for (int b = 0; b < nb; ++b) {
        const int4 c = ci[b];
        const float4 x0 = __ldg(xb + (size_t)(unsigned)c.x * 32), x1 = __ldg(xb + (size_t)(unsigned)c.y * 32);
        const float4 x2 = __ldg(xb + (size_t)(unsigned)c.z * 32), x3 = __ldg(xb + (size_t)(unsigned)c.w * 32);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!ROWVAL) v = vi[b];
        acc_row<ROWVAL>(a0, a1, v.x, x0);
        acc_row<ROWVAL>(a0, a1, v.y, x1);
        acc_row<ROWVAL>(a0, a1, v.z, x2);
        acc_row<ROWVAL>(a0, a1, v.w, x3);
      }
      const int rem = len & 3;                       // 0..3 entries left: one predicated batch (its padding slots are not read as rows)
      if (rem > 0) {
        const int4 c = ci[nb];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!ROWVAL) v = vi[nb];
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 x0 = z, x1 = z, x2 = z;
        x0 = __ldg(xb + (size_t)(unsigned)c.x * 32);
        if (rem > 1) x1 = __ldg(xb + (size_t)(unsigned)c.y * 32);
        if (rem > 2) x2 = __ldg(xb + (size_t)(unsigned)c.z * 32);
        acc_row<ROWVAL>(a0, a1, v.x, x0);
        if (rem > 1) acc_row<ROWVAL>(a0, a1, v.y, x1);
        if (rem > 2) acc_row<ROWVAL>(a0, a1, v.z, x2);
      }
      if (ROWVAL) {
        const float rv = __ldg(row_vals + row);
        a0.x *= rv; a0.y *= rv; a1.x *= rv; a1.y *= rv;
      }
      one[0] = make_float4(a0.x, a0.y, a1.x, a1.y);
    } else {                                         // rare: indices straight from global memory (shuffle-broadcast path)
      one[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      gather_range<32, 1, 4>(one, e0, e0 + len, len, 0, lane, colidx, vals, reinterpret_cast<const float*>(xb), 128);
    }
    epilogue<1>(ep, row, lane * 4, one, Y, 128);
  }
}

// chunks of the hub rows with the same lean loop: one warp per 256-entry chunk, its indices (and values) staged in the
// warp's slice of shared memory by one coalesced load of 8 entries per lane -> partial[chunk][128]
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_spmm_heavy2(int n_heavy, int n_chunks, const int32_t* __restrict__ heavy_rows,
                                                           const int32_t* __restrict__ chunk_ptr, const int32_t* __restrict__ rowptr,
                                                           const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                                           const float4* __restrict__ X4, float4* __restrict__ partial,
                                                           const int64_t row_begin, const int64_t row_end /*only hubs in this row range*/) {
  static_assert(GODE_HEAVY_CHUNK == 256, "eight entries per lane");
  __shared__ __align__(16) int s_idx[8][GODE_HEAVY_CHUNK];
  __shared__ __align__(16) float s_val[8][GODE_HEAVY_CHUNK];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int chunk = blockIdx.x * 8 + w;
  if (chunk >= n_chunks) return;
  // the hub row this chunk belongs to = last one whose first chunk is <= chunk: a 32-way search (the lanes probe 32 pivots
  // at once: 3 dependent loads for 30 000 hubs instead of the 15 of a binary search, all of them ahead of any useful work)
  int lo = 0, hi = n_heavy;
  while (hi - lo > 1) {
    const int step = (hi - lo + 31) >> 5;
    const int p = lo + lane * step;
    const bool le = p < hi && __ldg(chunk_ptr + p) <= chunk;
    const int k = __popc(__ballot_sync(0xffffffffu, le)) - 1;   // chunk_ptr is sorted and chunk_ptr[lo] <= chunk: lanes 0..k say yes
    lo += k * step;
    hi = min(lo + step, hi);
  }
  const int row = __ldg(heavy_rows + lo);
  if (row < row_begin || row >= row_end) return;    // (whole warp) a row-range call: the other ranges' hubs belong to other calls
  const int r0 = __ldg(rowptr + row), r1 = __ldg(rowptr + row + 1);
  const int e0 = r0 + (chunk - __ldg(chunk_ptr + lo)) * GODE_HEAVY_CHUNK;
  const int cnt = min(r1 - e0, GODE_HEAVY_CHUNK);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = k * 32 + lane;
    if (i < cnt) {
      s_idx[w][i] = __ldcs(colidx + e0 + i);
      s_val[w][i] = __ldcs(vals + e0 + i);
    }
  }
  __syncwarp();
  const float4* __restrict__ xb = X4 + lane;
  float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
  int j = 0;
#pragma unroll 1
  for (; j + 4 <= cnt; j += 4) {
    const int4 c = *reinterpret_cast<const int4*>(&s_idx[w][j]);      // broadcast LDS.128: one wavefront per four neighbours
    const float4 v = *reinterpret_cast<const float4*>(&s_val[w][j]);
    const float4 x0 = __ldg(xb + (size_t)(unsigned)c.x * 32), x1 = __ldg(xb + (size_t)(unsigned)c.y * 32);
    const float4 x2 = __ldg(xb + (size_t)(unsigned)c.z * 32), x3 = __ldg(xb + (size_t)(unsigned)c.w * 32);
    acc_row<false>(a0, a1, v.x, x0);
    acc_row<false>(a0, a1, v.y, x1);
    acc_row<false>(a0, a1, v.z, x2);
    acc_row<false>(a0, a1, v.w, x3);
  }
  for (; j < cnt; ++j) acc_row<false>(a0, a1, s_val[w][j], __ldg(xb + (size_t)(unsigned)s_idx[w][j] * 32));
  partial[(size_t)chunk * 32 + lane] = make_float4(a0.x, a0.y, a1.x, a1.y);
}


// L1-locality variant: ONE 1024-thread CTA per SM walks a contiguous tile of TILE_ROWS rows, so the neighbour
// rows its 32 warps gather (in a locality-ordered graph: a band around the tile) stay resident in that SM's L1
// and are served at L1 bandwidth (128 B/clk/SM) instead of L2 bandwidth (~42 B/clk/SM).
constexpr int TILE_WARPS = 32;
constexpr int TILE_ROUNDS = 4;   // rows per sub-warp

template <int LPR, int VPL>
__global__ void __launch_bounds__(TILE_WARPS * 32, 1) k_spmm_tile(int64_t n_rows, const int32_t* __restrict__ rowptr,
                                                                 const int32_t* __restrict__ colidx,
                                                                 const float* __restrict__ vals, const float* __restrict__ X,
                                                                 int64_t ldx, float* __restrict__ Y, int64_t ldy,
                                                                 const gode_spmm_epilogue_t ep) {
  constexpr int RPW = 32 / LPR;
  constexpr int ROWS_PER_ROUND = TILE_WARPS * RPW;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int sub = lane / LPR, sl = lane % LPR;
  const int col0 = sl * VPL * 4;
  const int64_t tile0 = blockIdx.x * (int64_t)(ROWS_PER_ROUND * TILE_ROUNDS);
#pragma unroll 1
  for (int r = 0; r < TILE_ROUNDS; ++r) {
    const int64_t row = tile0 + (int64_t)r * ROWS_PER_ROUND + w * RPW + sub;
    const bool valid = row < n_rows;
    int e0 = 0, e1 = 0;
    if (valid) {
      e0 = __ldg(rowptr + row);
      e1 = __ldg(rowptr + row + 1);
    }
    const bool heavy = (e1 - e0) > GODE_HEAVY_ROW;
    if (heavy) e1 = e0;
    const int maxlen = warp_max_over_subs<LPR>(e1 - e0);
    float4 acc[VPL];
#pragma unroll
    for (int u = 0; u < VPL; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    gather_range<LPR, VPL, 4>(acc, e0, e1, maxlen, sub, sl, colidx, vals, X + col0, ldx);
    if (valid && !heavy) epilogue<VPL>(ep, row, col0, acc, Y, ldy);
  }
}


// ------------------------------------------------------------------------------------------------
// Persistent-warp gather (default).  ncu on the one-row-per-warp kernel above (N=2M, 39M entries, d=128):
// 26 % of the stall samples sit on the rowptr -> (col,val) dependency chain at the head of every row and the
// L1 data pipe is 60 % busy (a 512 B row costs 4 wavefronts, each broadcast shuffle one more).  Here
//   * a fixed grid of warps strides over the rows; the (col,val) batch of the NEXT row and the rowptr pair of
//     the row after it are already in flight while the current row's neighbour rows are gathered;
//   * there is a single predicated code path (no slow tail): U loads are always issued back to back;
//   * ROWVAL: when the matrix is row-constant (gode_csr_t.row_vals, the reference's D^-1(A+I)) the values
//     stream and its broadcast shuffle disappear: acc = row_val * sum_e X[col_e].
// ------------------------------------------------------------------------------------------------
template <int LPR, int VPL, int U, bool ROWVAL>
__device__ __forceinline__ void consume_batch(float4 (&acc)[VPL], int c, float v, int cnt, int maxcnt, int sub,
                                              const float* __restrict__ xl, int64_t ldx) {
  constexpr int UU = U < LPR ? U : LPR;
#pragma unroll
  for (int j = 0; j < LPR; j += UU) {
    if (j >= maxcnt) break;  // warp-uniform
    int cj[UU];
    float vj[UU];
#pragma unroll
    for (int q = 0; q < UU; ++q) {
      cj[q] = __shfl_sync(0xffffffffu, c, sub * LPR + j + q);
      vj[q] = ROWVAL ? 1.f : __shfl_sync(0xffffffffu, v, sub * LPR + j + q);
    }
    float4 x[UU][VPL];
#pragma unroll
    for (int q = 0; q < UU; ++q) {
      const bool on = j + q < cnt;
#pragma unroll
      for (int u = 0; u < VPL; ++u) {
        x[q][u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on) x[q][u] = ld_ro4(xl + (int64_t)cj[q] * ldx + u * 4);
      }
    }
#pragma unroll
    for (int q = 0; q < UU; ++q)
#pragma unroll
      for (int u = 0; u < VPL; ++u) {
        if (ROWVAL) {
          acc[u].x += x[q][u].x; acc[u].y += x[q][u].y; acc[u].z += x[q][u].z; acc[u].w += x[q][u].w;
        } else {
          acc[u].x += vj[q] * x[q][u].x; acc[u].y += vj[q] * x[q][u].y;
          acc[u].z += vj[q] * x[q][u].z; acc[u].w += vj[q] * x[q][u].w;
        }
      }
  }
}

template <int LPR, int VPL, int U, int MINB, bool ROWVAL>
__global__ void __launch_bounds__(256, MINB) k_spmm_pw(int64_t n_rows, const int32_t* __restrict__ rowptr,
                                                       const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                                       const float* __restrict__ row_vals, const float* __restrict__ X,
                                                       int64_t ldx, float* __restrict__ Y, int64_t ldy,
                                                       const gode_spmm_epilogue_t ep) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const int col0 = sl * VPL * 4;
  const float* __restrict__ xl = X + col0;
  const int64_t n_units = (n_rows + RPW - 1) / RPW;
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5);
  int64_t unit = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (unit >= n_units) return;

  auto load_ptr = [&](int64_t u, int& e0, int& e1, float& rv, bool& heavy) {
    const int64_t r = u * RPW + sub;
    e0 = 0; e1 = 0; rv = 0.f; heavy = false;
    if (u < n_units && r < n_rows) {
      e0 = __ldg(rowptr + r);
      e1 = __ldg(rowptr + r + 1);
      if (ROWVAL) rv = __ldg(row_vals + r);
    }
  };
  auto load_cv = [&](int e0, int e1, int off, int& c, float& v) {
    c = 0; v = 0.f;
    const int e = e0 + off + sl;
    if (e < e1) {
      c = __ldcs(colidx + e);
      if (!ROWVAL) v = __ldcs(vals + e);
    }
  };

  // pipeline registers: current row (e0,e1,rv,c,v), next row (pointers + first batch), row after next (pointers)
  int e0, e1, e0n, e1n, e0nn, e1nn, c, cn;
  float rv, rvn, rvnn, v, vn;
  bool hv, hvn, hvnn;
  load_ptr(unit, e0, e1, rv, hv);
  load_ptr(unit + stride, e0n, e1n, rvn, hvn);
  if (e1 - e0 > GODE_HEAVY_ROW) { hv = true; e1 = e0; }
  load_cv(e0, e1, 0, c, v);

  for (; unit < n_units; unit += stride) {
    // prefetch: next row's first batch, and the pointers of the row after next
    if (e1n - e0n > GODE_HEAVY_ROW) { hvn = true; e1n = e0n; }
    load_cv(e0n, e1n, 0, cn, vn);
    load_ptr(unit + 2 * stride, e0nn, e1nn, rvnn, hvnn);

    float4 acc[VPL];
#pragma unroll
    for (int u = 0; u < VPL; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int len = e1 - e0;
    const int maxlen = warp_max_over_subs<LPR>(len);
    consume_batch<LPR, VPL, U, ROWVAL>(acc, c, v, min(len, LPR), min(maxlen, LPR), sub, xl, ldx);
    for (int off = LPR; off < maxlen; off += LPR) {   // rows longer than one batch (uncommon)
      load_cv(e0, e1, off, c, v);
      consume_batch<LPR, VPL, U, ROWVAL>(acc, c, v, min(len - off, LPR), min(maxlen - off, LPR), sub, xl, ldx);
    }
    const int64_t row = unit * RPW + sub;
    if (row < n_rows && !hv) {
      if (ROWVAL) {
#pragma unroll
        for (int u = 0; u < VPL; ++u) { acc[u].x *= rv; acc[u].y *= rv; acc[u].z *= rv; acc[u].w *= rv; }
      }
      epilogue<VPL>(ep, row, col0, acc, Y, ldy);
    }
    e0 = e0n; e1 = e1n; rv = rvn; hv = hvn; c = cn; v = vn;
    e0n = e0nn; e1n = e1nn; rvn = rvnn; hvn = hvnn;
  }
}

template <int LPR, int VPL, int U, int MINB>
static int launch_pw(const gode_csr_t& A, const float* X, int64_t ldx, float* Y, int64_t ldy,
                     const gode_spmm_epilogue_t& ep, cudaStream_t st) {
  constexpr int RPW = 32 / LPR;
  const int64_t n_units = (A.n_rows + RPW - 1) / RPW;
  int64_t blocks = (n_units + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * MINB;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) return GODE_OK;
  if (A.row_vals)
    k_spmm_pw<LPR, VPL, U, MINB, true><<<static_cast<unsigned>(blocks), 256, 0, st>>>(A.n_rows, A.rowptr, A.colidx, A.vals,
                                                                                       A.row_vals, X, ldx, Y, ldy, ep);
  else
    k_spmm_pw<LPR, VPL, U, MINB, false><<<static_cast<unsigned>(blocks), 256, 0, st>>>(A.n_rows, A.rowptr, A.colidx, A.vals,
                                                                                        nullptr, X, ldx, Y, ldy, ep);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

// one sub-warp per chunk of a heavy row -> partial[chunk][d]
template <int LPR, int VPL, int MINB = 1, int UN = UNR>
__global__ void __launch_bounds__(256, MINB) k_spmm_heavy_partial(int n_heavy, int n_chunks, const int32_t* __restrict__ heavy_rows,
                                                            const int32_t* __restrict__ chunk_ptr,
                                                            const int32_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                                            const float* __restrict__ X, int64_t ldx,
                                                            float* __restrict__ partial) {
  constexpr int RPW = 32 / LPR;
  constexpr int D = LPR * VPL * 4;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const int chunk = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + sub;
  int e0 = 0, e1 = 0;
  if (chunk < n_chunks) {
    int lo = 0, hi = n_heavy - 1;  // last heavy row whose first chunk is <= chunk
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(chunk_ptr + mid) <= chunk) lo = mid; else hi = mid - 1;
    }
    const int row = __ldg(heavy_rows + lo);
    const int r0 = __ldg(rowptr + row), r1 = __ldg(rowptr + row + 1);
    e0 = r0 + (chunk - __ldg(chunk_ptr + lo)) * GODE_HEAVY_CHUNK;
    e1 = min(r1, e0 + GODE_HEAVY_CHUNK);
  }
  const int maxlen = warp_max_over_subs<LPR>(e1 - e0);
  const int col0 = sl * VPL * 4;
  float4 acc[VPL];
#pragma unroll
  for (int u = 0; u < VPL; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  gather_range<LPR, VPL, UN>(acc, e0, e1, maxlen, sub, sl, colidx, vals, X + col0, ldx);
  if (chunk < n_chunks) {
#pragma unroll
    for (int u = 0; u < VPL; ++u) *reinterpret_cast<float4*>(partial + (int64_t)chunk * D + col0 + u * 4) = acc[u];
  }
}

// one sub-warp per heavy row: add its chunks in order, then the row epilogue
template <int LPR, int VPL>
__global__ void __launch_bounds__(256) k_spmm_heavy_finish(int n_heavy, const int32_t* __restrict__ heavy_rows,
                                                           const int32_t* __restrict__ chunk_ptr,
                                                           const float* __restrict__ partial, float* __restrict__ Y,
                                                           int64_t ldy, const gode_spmm_epilogue_t ep,
                                                           const int64_t row_begin = 0, const int64_t row_end = INT64_MAX) {
  constexpr int RPW = 32 / LPR;
  constexpr int D = LPR * VPL * 4;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const int h = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + sub;
  if (h >= n_heavy) return;
  if (heavy_rows[h] < row_begin || heavy_rows[h] >= row_end) return;
  const int c0 = chunk_ptr[h], c1 = chunk_ptr[h + 1];
  const int col0 = sl * VPL * 4;
  float4 acc[VPL];
#pragma unroll
  for (int u = 0; u < VPL; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = c0; c < c1; ++c) {
#pragma unroll
    for (int u = 0; u < VPL; ++u) {
      const float4 p = *reinterpret_cast<const float4*>(partial + (int64_t)c * D + col0 + u * 4);
      acc[u].x += p.x; acc[u].y += p.y; acc[u].z += p.z; acc[u].w += p.w;
    }
  }
  epilogue<VPL>(ep, heavy_rows[h], col0, acc, Y, ldy);
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
      if (ep.acc_in) v += ep.acc_in[row * ldy + c];
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
      if (ep.second.out) {
        float t = 0.f;
        for (int j = 0; j < ep.n_prev; ++j) t += ep.second.coef[j] * ep.kprev[j][o];
        t += ep.second.coef_self * v;
        ep.second.out[o] = ep.y0[o] + t;
      }
      if (ep.gp_out) {
        const float ms = ep.gp_row_scale ? ep.mask_scale * ep.gp_row_scale[row] : ep.mask_scale;
        ep.gp_out[o] = v > 0.f ? ms * ep.mask_src[o] : 0.f;
      }
    }
  }
}

template <int LPR, int VPL>
static int launch_vec(const gode_csr_t& A, const float* X, int64_t ldx, float* Y, int64_t ldy,
                      const gode_spmm_epilogue_t& ep, float* ws, cudaStream_t st, int64_t row_begin = 0, int64_t row_end = -1) {
  if (row_end < 0) row_end = A.n_rows;
  const bool sub = row_begin != 0 || row_end != A.n_rows;
  constexpr int RPW = 32 / LPR;
  constexpr int RPB = 8 * RPW;
  static const int variant = [] {
    const char* e = getenv("GODE_SPMM_VARIANT");
    // 8 (default where it applies: d = 128, contiguous rows): k_spmm_t2, row tiles staged in shared memory + lean inner loop.
    //    Measured at N = 10 M (bare A gather / A^T gather / gather with the rk4 stage-4 epilogue, ms): 6.72 / 8.06 / 9.69
    //    against 8.85 / 9.42 / 11.49 for variant 0 (profiles/r02_gather.md).
    // 7: k_spmm_ts (tile staging without the lean loop), 6: k_spmm_sb (shared-memory index broadcast) -- kept for the record.
    // 0 (every other width): one row per sub-warp, CTAs dispatched in row order -- the hardware scheduler keeps the rows in
    //    flight inside a narrow id window, which is what lets L2 hold the gather band of a locality-ordered graph.
    // 5/4: persistent warps (U=8 x 3 CTAs/SM | U=4 x 4 CTAs/SM): faster at N <= 2M, but the warps drift apart on a
    //    power-law graph (ncu at N=10M: L2 hit rate 26 %, 52 GB of DRAM reads per launch -> DRAM-bound, 7 % slower).
    // 3: one 1024-thread CTA per SM on a row tile.
    return e ? atoi(e) : 8;
  }();
  static const int minb_sb = [] {
    const char* e = getenv("GODE_SPMM_MINB");
    return e ? atoi(e) : 8;
  }();
  static const int unr_sb = [] {
    const char* e = getenv("GODE_SPMM_UNR");
    return e ? atoi(e) : 4;
  }();
  static const int use_rowval = [] {
    const char* e = getenv("GODE_SPMM_ROWVAL");   // 1 (default): skip the values stream of a row-constant matrix
    return e ? atoi(e) : 1;
  }();
  if (LPR == 32 && variant == 6) {
    if constexpr (LPR == 32) {
      // k_spmm_sb: indices broadcast through shared memory (see the kernel's header)
      static const int prefetch_sb = [] {
        const char* e = getenv("GODE_SPMM_PREFETCH");
        return e ? atoi(e) : 1;
      }();
      const bool rv = use_rowval && A.row_vals != nullptr;
      if (A.n_rows > 0) {
        unsigned grid = static_cast<unsigned>((A.n_rows + 7) / 8);
#define GODE_SB_LAUNCH(MB, UU)                                                                                          \
  do {                                                                                                                  \
    if (rv) k_spmm_sb<VPL, MB, UU, true><<<grid, 256, 0, st>>>(A.n_rows, A.rowptr, A.colidx, A.vals, A.row_vals, X, ldx, Y, ldy, ep, prefetch_sb); \
    else k_spmm_sb<VPL, MB, UU, false><<<grid, 256, 0, st>>>(A.n_rows, A.rowptr, A.colidx, A.vals, nullptr, X, ldx, Y, ldy, ep, prefetch_sb);    \
  } while (0)
        if (VPL > 1 || (minb_sb == 4 && unr_sb == 8)) GODE_SB_LAUNCH(4, 8);   // d = 256: two float4 per lane need the registers
        else if (minb_sb == 6 && unr_sb == 8) GODE_SB_LAUNCH(6, 8);
        else if (minb_sb == 6 && unr_sb == 4) GODE_SB_LAUNCH(6, 4);
        else GODE_SB_LAUNCH(8, 4);
#undef GODE_SB_LAUNCH
        GODE_LAUNCH_CHECK();
      }
      if (A.n_heavy > 0) {
        unsigned g1 = static_cast<unsigned>((A.n_chunks + 7) / 8);
        k_spmm_sb_heavy<VPL, 8><<<g1, 256, 0, st>>>(A.n_heavy, A.n_chunks, A.heavy_rows, A.heavy_chunk_ptr, A.rowptr, A.colidx,
                                                    A.vals, X, ldx, ws);
        GODE_LAUNCH_CHECK();
        unsigned g2 = static_cast<unsigned>((A.n_heavy + RPB - 1) / RPB);
        k_spmm_heavy_finish<LPR, VPL><<<g2, 256, 0, st>>>(A.n_heavy, A.heavy_rows, A.heavy_chunk_ptr, ws, Y, ldy, ep);
        GODE_LAUNCH_CHECK();
      }
      return GODE_OK;
    }
  }
  if ((variant == 8 || sub) && LPR == 32 && VPL == 1 && ldx == 128 && (ldy == 128 || !Y)) {
    if constexpr (LPR == 32 && VPL == 1) {
      static const int prefetch_t2 = [] {
        const char* e = getenv("GODE_SPMM_PREFETCH");   // default off here: the L2 prefetch of the epilogue operands costs LSU
        return e ? atoi(e) : 0;                         // wavefronts, which is what this kernel is bound by (9.94 vs 10.19 ms)
      }();
      static const int t2_rows = [] {
        const char* e = getenv("GODE_SPMM_TS_ROWS");
        return e ? atoi(e) : 32;
      }();
      const bool rv = use_rowval && A.row_vals != nullptr;
      const float4* X4 = reinterpret_cast<const float4*>(X);
      static const int use_sched = [] {
        const char* e = getenv("GODE_SPMM_SCHED");   // 1 (default): hub chunks are CTAs of the tile grid, in the plan's work order
        return e ? atoi(e) : 1;                      // (gode_csr_t.tile_sched); 0: separate hub kernel after the tiles
      }();
      static const int t3 = [] {
        const char* e = getenv("GODE_SPMM_T3");      // 1: aligned staging + LDS.128 index reads (k_spmm_t3).  Default off: measured
        return e ? atoi(e) : 0;                      // 6.68 / 7.93 / 9.59 ms against k_spmm_t2's 6.47 / 7.59 / 9.04 (bare A, A^T, fused epilogue)
      }();
      HubArgs hub;
      hub.sched = (use_sched && A.n_heavy > 0 && A.tile_sched && A.n_tile_sched > 0 && t2_rows == 32 && !t3 && !sub) ? A.tile_sched
                                                                                                                   : nullptr;
      hub.rows = A.heavy_rows;
      hub.chunk_ptr = A.heavy_chunk_ptr;
      hub.n_heavy = A.n_heavy;
      hub.n_chunks = A.n_chunks;
      hub.partial = reinterpret_cast<float4*>(ws);
      if (row_end > row_begin) {
#define GODE_T2_LAUNCH(R, MB)                                                                                           \
  do {                                                                                                                  \
    unsigned grid = hub.sched ? static_cast<unsigned>(A.n_tile_sched) : static_cast<unsigned>((row_end - row_begin + R - 1) / R); \
    if (rv) k_spmm_t2<R, MB, true><<<grid, 256, 0, st>>>(row_end, A.rowptr, A.colidx, A.vals, A.row_vals, X4, Y, ep, prefetch_t2, hub, row_begin); \
    else k_spmm_t2<R, MB, false><<<grid, 256, 0, st>>>(row_end, A.rowptr, A.colidx, A.vals, nullptr, X4, Y, ep, prefetch_t2, hub, row_begin);    \
  } while (0)
        if (t2_rows == 32 && t3 && !sub) {
          unsigned grid = static_cast<unsigned>((A.n_rows + 31) / 32);
#define GODE_T3_LAUNCH(MB)                                                                                              \
  do {                                                                                                                  \
    if (rv) k_spmm_t3<MB, true><<<grid, 256, 0, st>>>(A.n_rows, A.rowptr, A.colidx, A.vals, A.row_vals, X4, Y, ep, prefetch_t2); \
    else k_spmm_t3<MB, false><<<grid, 256, 0, st>>>(A.n_rows, A.rowptr, A.colidx, A.vals, nullptr, X4, Y, ep, prefetch_t2);    \
  } while (0)
          if (minb_sb == 6) GODE_T3_LAUNCH(6); else if (minb_sb == 7) GODE_T3_LAUNCH(7); else GODE_T3_LAUNCH(8);
#undef GODE_T3_LAUNCH
        } else if (t2_rows == 32) {
          if (minb_sb == 6) GODE_T2_LAUNCH(32, 6); else GODE_T2_LAUNCH(32, 8);
        } else if (t2_rows == 128) {
          if (minb_sb == 6) GODE_T2_LAUNCH(128, 6); else GODE_T2_LAUNCH(128, 8);
        } else {
          if (minb_sb == 6) GODE_T2_LAUNCH(64, 6);
          else if (minb_sb == 7) GODE_T2_LAUNCH(64, 7);
          else if (minb_sb == 5) GODE_T2_LAUNCH(64, 5);
          else if (minb_sb == 4) GODE_T2_LAUNCH(64, 4);
          else GODE_T2_LAUNCH(64, 8);
        }
#undef GODE_T2_LAUNCH
        GODE_LAUNCH_CHECK();
      }
      if (A.n_heavy > 0) {
        if (!hub.sched) {
          unsigned g1 = static_cast<unsigned>((A.n_chunks + 7) / 8);
          k_spmm_heavy2<6><<<g1, 256, 0, st>>>(A.n_heavy, A.n_chunks, A.heavy_rows, A.heavy_chunk_ptr, A.rowptr, A.colidx, A.vals,
                                               X4, reinterpret_cast<float4*>(ws), row_begin, row_end);
          GODE_LAUNCH_CHECK();
        }
        unsigned g2 = static_cast<unsigned>((A.n_heavy + RPB - 1) / RPB);
        k_spmm_heavy_finish<LPR, VPL><<<g2, 256, 0, st>>>(A.n_heavy, A.heavy_rows, A.heavy_chunk_ptr, ws, Y, ldy, ep, row_begin, row_end);
        GODE_LAUNCH_CHECK();
      }
      return GODE_OK;
    }
  }
  if (sub) {
    set_error("spmm: a row-range call needs d = 128 with contiguous rows (the k_spmm_t2 path)");
    return GODE_EINVAL;
  }
  if (variant == 7 && LPR == 32 && VPL == 1) {
    if constexpr (LPR == 32 && VPL == 1) {
      static const int prefetch_ts = [] {
        const char* e = getenv("GODE_SPMM_PREFETCH");
        return e ? atoi(e) : 1;
      }();
      const bool rv = use_rowval && A.row_vals != nullptr;
      static const int ts_rows = [] {
        const char* e = getenv("GODE_SPMM_TS_ROWS");
        return e ? atoi(e) : 64;
      }();
      static const int pf_dist = [] {
        const char* e = getenv("GODE_SPMM_PFDIST");   // tiles ahead whose neighbour rows are requested from L2 (0 = off)
        return e ? atoi(e) : 0;
      }();
      if (A.n_rows > 0) {
#define GODE_TS_LAUNCH(R, MB, UU)                                                                                       \
  do {                                                                                                                  \
    unsigned grid = static_cast<unsigned>((A.n_rows + R - 1) / R);                                                      \
    if (rv) k_spmm_ts<R, MB, UU, true><<<grid, 256, 0, st>>>(A.n_rows, A.rowptr, A.colidx, A.vals, A.row_vals, X, ldx, Y, ldy, ep, prefetch_ts, pf_dist); \
    else k_spmm_ts<R, MB, UU, false><<<grid, 256, 0, st>>>(A.n_rows, A.rowptr, A.colidx, A.vals, nullptr, X, ldx, Y, ldy, ep, prefetch_ts, pf_dist);    \
  } while (0)
        if (ts_rows == 16) {
          if (minb_sb == 6) GODE_TS_LAUNCH(16, 6, 4); else GODE_TS_LAUNCH(16, 8, 4);
        } else if (ts_rows == 32) {
          if (minb_sb == 6) GODE_TS_LAUNCH(32, 6, 4); else GODE_TS_LAUNCH(32, 8, 4);
        } else if (ts_rows == 128) {
          if (minb_sb == 6) GODE_TS_LAUNCH(128, 6, 4); else GODE_TS_LAUNCH(128, 8, 4);
        } else {
          if (minb_sb == 6 && unr_sb == 8) GODE_TS_LAUNCH(64, 6, 8);
          else if (minb_sb == 6) GODE_TS_LAUNCH(64, 6, 4);
          else if (minb_sb == 5) GODE_TS_LAUNCH(64, 5, 4);
          else if (minb_sb == 7) GODE_TS_LAUNCH(64, 7, 4);
          else GODE_TS_LAUNCH(64, 8, 4);
        }
#undef GODE_TS_LAUNCH
        GODE_LAUNCH_CHECK();
      }
      if (A.n_heavy > 0) {
        unsigned g1 = static_cast<unsigned>((A.n_chunks + RPB - 1) / RPB);
        k_spmm_heavy_partial<LPR, VPL><<<g1, 256, 0, st>>>(A.n_heavy, A.n_chunks, A.heavy_rows, A.heavy_chunk_ptr, A.rowptr,
                                                          A.colidx, A.vals, X, ldx, ws);
        GODE_LAUNCH_CHECK();
        unsigned g2 = static_cast<unsigned>((A.n_heavy + RPB - 1) / RPB);
        k_spmm_heavy_finish<LPR, VPL><<<g2, 256, 0, st>>>(A.n_heavy, A.heavy_rows, A.heavy_chunk_ptr, ws, Y, ldy, ep);
        GODE_LAUNCH_CHECK();
      }
      return GODE_OK;
    }
  }
  if (A.n_rows > 0 && (variant == 5 || variant == 4)) {
    int rc = variant == 5 ? launch_pw<LPR, VPL, 8, 3>(A, X, ldx, Y, ldy, ep, st) : launch_pw<LPR, VPL, 4, 4>(A, X, ldx, Y, ldy, ep, st);
    if (rc) return rc;
  } else if (A.n_rows > 0 && variant == 3) {
    constexpr int TR = TILE_WARPS * RPW * TILE_ROUNDS;
    unsigned grid = static_cast<unsigned>((A.n_rows + TR - 1) / TR);
    k_spmm_tile<LPR, VPL><<<grid, TILE_WARPS * 32, 0, st>>>(A.n_rows, A.rowptr, A.colidx, A.vals, X, ldx, Y, ldy, ep);
    GODE_LAUNCH_CHECK();
  } else if (A.n_rows > 0) {
    unsigned grid = static_cast<unsigned>((A.n_rows + RPB - 1) / RPB);
    static const int prefetch = [] {
      const char* e = getenv("GODE_SPMM_PREFETCH");   // 1 (default): L2-prefetch the epilogue operands before the gather
      return e ? atoi(e) : 1;
    }();
    // Resident CTAs per SM (register cap) x neighbour rows in flight per lane.  The compiler's own choice is 64
    // registers = 4 CTAs; measured at N = 10 M (ms per bare A^T gather / per gather with epilogue):
    //   4 CTAs x 8 rows: 12.63 / 15.44   6 x 8: 11.21 / 13.31   6 x 4: 10.47 / 12.51   8 x 4: 9.95 / 11.77  <- default
    // (GODE_SPMM_MINB, GODE_SPMM_UNR): full occupancy with short batches wins -- the gather is a latency chain per warp
    // (rowptr -> colidx -> batches of neighbour rows), and 64 resident warps hide it better than deeper batches do.
    static const int minb = [] {
      const char* e = getenv("GODE_SPMM_MINB");
      return e ? atoi(e) : 8;
    }();
    static const int unr = [] {
      const char* e = getenv("GODE_SPMM_UNR");
      return e ? atoi(e) : 4;
    }();
#define GODE_VEC_LAUNCH(MB, UU)                                                                                     \
  k_spmm_vec<LPR, VPL, MB, UU><<<grid, 256, 0, st>>>(A.n_rows, A.rowptr, A.colidx, A.vals, X, ldx, Y, ldy, ep, prefetch)
    if constexpr (LPR == 32 && VPL == 1) {
      if (minb == 5 && unr == 8) GODE_VEC_LAUNCH(5, 8);
      else if (minb == 6 && unr == 8) GODE_VEC_LAUNCH(6, 8);
      else if (minb == 7 && unr == 8) GODE_VEC_LAUNCH(7, 8);
      else if (minb == 8 && unr == 8) GODE_VEC_LAUNCH(8, 8);
      else if (minb == 6 && unr == 4) GODE_VEC_LAUNCH(6, 4);
      else if (minb == 8 && unr == 4) GODE_VEC_LAUNCH(8, 4);
      else if (minb == 8 && unr == 2) GODE_VEC_LAUNCH(8, 2);
      else GODE_VEC_LAUNCH(1, 8);
    } else {
      GODE_VEC_LAUNCH(1, UNR);
    }
#undef GODE_VEC_LAUNCH
    GODE_LAUNCH_CHECK();
  }
  if (A.n_heavy > 0) {
    unsigned g1 = static_cast<unsigned>((A.n_chunks + RPB - 1) / RPB);
    // chunks are full 256-entry ranges: deep batches (8 rows in flight, the compiler's 44 registers) suit them; the
    // 8-CTA / 4-row point of the main kernel measured no better here (10.19 vs 9.95 ms per bare gather in total)
    k_spmm_heavy_partial<LPR, VPL><<<g1, 256, 0, st>>>(A.n_heavy, A.n_chunks, A.heavy_rows, A.heavy_chunk_ptr, A.rowptr,
                                                      A.colidx, A.vals, X, ldx, ws);
    GODE_LAUNCH_CHECK();
    unsigned g2 = static_cast<unsigned>((A.n_heavy + RPB - 1) / RPB);
    k_spmm_heavy_finish<LPR, VPL><<<g2, 256, 0, st>>>(A.n_heavy, A.heavy_rows, A.heavy_chunk_ptr, ws, Y, ldy, ep);
    GODE_LAUNCH_CHECK();
  }
  return GODE_OK;
}


// ------------------------------------------------------------------------------------------------
// Bulk-copy (TMA engine) gather pipeline for wide rows (d = 128 / 256).
//
// The register-staged kernel above can keep only UNR neighbour rows per lane in flight, and every warp pays
// the rowptr -> colidx -> gather dependency chain once per row: measured 1.1 ms for 1M rows / 20M entries
// (13 % of the HBM roofline) -- latency-bound, not bandwidth-bound.  Here every lane issues ONE
// cp.async.bulk (global -> shared, a whole 512 B / 1 KB neighbour row, completion counted on an mbarrier),
// so a warp has up to 2 x 32 rows in flight without holding a register for them; the (col,val) pairs of the
// batch after next are prefetched into registers while the next batch's rows are in flight and the current
// batch is consumed from shared memory.  One warp walks a tile of R consecutive rows; batches never cross a
// row, so the accumulator is a register float4 per lane and the fused epilogue is unchanged.
//   smem per warp: 2 buffers x 32 slots x d x 4 B (32 KB at d=128)  ->  NW warps per CTA, one CTA per SM.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct Batch {
  int lr;     // local row (or chunk) index in the warp's tile
  int off;    // offset of the batch in its row
  int e;      // first entry
  int cnt;    // entries in the batch (0 for an empty row)
  bool last;  // final batch of its row
  bool valid;
};

constexpr int BULK_NW = 6;   // warps per CTA
constexpr int BULK_R = 16;   // rows (or heavy-row chunks) per warp

// CHUNKS = false: rows r0.. of the matrix (heavy rows skipped), epilogue per row.
// CHUNKS = true : chunks of the heavy rows, partial sums to `partial`.
// ASYNC = 0: one cp.async.bulk (TMA engine) per neighbour row, issued by one lane, mbarrier completion.
// ASYNC = 1: one warp-wide cp.async (LDGSTS, 16 B per lane) per neighbour row, commit/wait groups.
// Measured (N=1M, 20M entries, d=128): the TMA engine sustains only about one 512 B request per ~55 cycles
// per SM (2.6 TB/s chip-wide), so per-row bulk copies lose to LDGSTS; both are kept for the record.
template <int VPL, bool CHUNKS, int ASYNC>
__global__ void __launch_bounds__(BULK_NW * 32, 1)
k_spmm_bulk(int64_t n_units, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
            const float* __restrict__ vals, const int32_t* __restrict__ heavy_rows, const int32_t* __restrict__ chunk_ptr,
            int n_heavy, const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy,
            float* __restrict__ partial, const gode_spmm_epilogue_t ep) {
  constexpr int D = 32 * VPL * 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ring = reinterpret_cast<float*>(smem_raw) + (size_t)warp * 2 * 32 * D;          // [2][32][D]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)BULK_NW * 2 * 32 * D * 4) + warp * 2;
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const int64_t u0 = (blockIdx.x * (int64_t)BULK_NW + warp) * BULK_R;
  const int nu = static_cast<int>(min((int64_t)BULK_R, n_units - u0));
  if (nu <= 0) return;

  // lane i < nu holds [rs, re) = entry range of local unit i, and whether its epilogue must be skipped
  int rs = 0, re = 0;
  bool skip = false;
  if (lane < nu) {
    if (!CHUNKS) {
      rs = __ldg(rowptr + u0 + lane);
      re = __ldg(rowptr + u0 + lane + 1);
      if (re - rs > GODE_HEAVY_ROW) { skip = true; re = rs; }
    } else {
      const int chunk = static_cast<int>(u0) + lane;
      int lo = 0, hi = n_heavy - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(chunk_ptr + mid) <= chunk) lo = mid; else hi = mid - 1;
      }
      const int row = __ldg(heavy_rows + lo);
      const int r0 = __ldg(rowptr + row), r1 = __ldg(rowptr + row + 1);
      rs = r0 + (chunk - __ldg(chunk_ptr + lo)) * GODE_HEAVY_CHUNK;
      re = min(r1, rs + GODE_HEAVY_CHUNK);
    }
  }

  auto make = [&](int lr, int off) {
    Batch b;
    b.lr = lr; b.off = off; b.valid = lr < nu;
    const int src = b.valid ? lr : 0;
    const int s = __shfl_sync(0xffffffffu, rs, src), e = __shfl_sync(0xffffffffu, re, src);
    b.e = s + off;
    const int rem = e - s - off;
    b.cnt = b.valid ? max(0, min(32, rem)) : 0;
    b.last = rem <= 32;
    return b;
  };
  auto advance = [&](const Batch& b) {
    if (!b.valid) return b;
    return b.last ? make(b.lr + 1, 0) : make(b.lr, b.off + 32);
  };
  auto load_cv = [&](const Batch& b, int& c, float& v) {
    c = 0; v = 0.f;
    if (lane < b.cnt) {
      c = __ldcs(colidx + b.e + lane);
      v = __ldcs(vals + b.e + lane);
    }
  };
  auto issue = [&](const Batch& b, int c, int buf) {
    if (ASYNC == 0) {
      if (b.cnt <= 0) return;
      if (lane == 0) mbar_expect_tx(&bars[buf], static_cast<uint32_t>(b.cnt) * D * 4u);
      __syncwarp();
      if (lane < b.cnt) bulk_g2s(ring + ((size_t)buf * 32 + lane) * D, X + (int64_t)c * ldx, D * 4u, &bars[buf]);
    } else {
      const uint32_t dst0 = smem_u32(ring + (size_t)buf * 32 * D + lane * VPL * 4);
      const float* src0 = X + lane * VPL * 4;
#pragma unroll 4
      for (int j = 0; j < b.cnt; ++j) {
        const int cj = __shfl_sync(0xffffffffu, c, j);
#pragma unroll
        for (int u = 0; u < VPL; ++u)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)(j * D + u * 4) * 4u),
                       "l"(src0 + (int64_t)cj * ldx + u * 4)
                       : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");   // one group per batch, empty batches included
    }
  };

  Batch b0 = make(0, 0), b1 = advance(b0), b2 = advance(b1);
  int c0, c1, c2 = 0;
  float v0, v1, v2 = 0.f;
  load_cv(b0, c0, v0);
  issue(b0, c0, 0);
  load_cv(b1, c1, v1);
  uint32_t phase[2] = {0u, 0u};
  float4 acc[VPL];
#pragma unroll
  for (int u = 0; u < VPL; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int col0 = lane * VPL * 4;

  for (int k = 0; b0.valid; ++k) {
    const int buf = k & 1;
    if (b1.valid) issue(b1, c1, buf ^ 1);
    if (b2.valid) load_cv(b2, c2, v2);
    if (ASYNC == 1) {
      // groups are committed in batch order; all but the newest (b1's) must have landed
      if (b1.valid) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
    }
    if (b0.cnt > 0) {
      if (ASYNC == 0) {
        mbar_wait(&bars[buf], phase[buf]);
        phase[buf] ^= 1u;
      }
      const float* rb = ring + (size_t)buf * 32 * D + col0;
      constexpr int USTRIDE = 4;
#pragma unroll 8
      for (int j = 0; j < b0.cnt; ++j) {
        const float vj = __shfl_sync(0xffffffffu, v0, j);
#pragma unroll
        for (int u = 0; u < VPL; ++u) {
          const float4 x = *reinterpret_cast<const float4*>(rb + (size_t)j * D + u * USTRIDE);
          acc[u].x += vj * x.x; acc[u].y += vj * x.y; acc[u].z += vj * x.z; acc[u].w += vj * x.w;
        }
      }
    }
    if (b0.last) {
      if (CHUNKS) {
#pragma unroll
        for (int u = 0; u < VPL; ++u)
          *reinterpret_cast<float4*>(partial + (u0 + b0.lr) * D + col0 + u * 4) = acc[u];
      } else {
        const bool sk = __shfl_sync(0xffffffffu, skip ? 1 : 0, b0.lr) != 0;
        if (!sk) epilogue<VPL>(ep, u0 + b0.lr, col0, acc, Y, ldy);
      }
#pragma unroll
      for (int u = 0; u < VPL; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    b0 = b1; b1 = b2; b2 = advance(b2);
    c0 = c1; v0 = v1; c1 = c2; v1 = v2;
  }
}

template <int VPL, int ASYNC>
static int launch_bulk(const gode_csr_t& A, const float* X, int64_t ldx, float* Y, int64_t ldy,
                       const gode_spmm_epilogue_t& ep, float* ws, cudaStream_t st) {
  constexpr int D = 32 * VPL * 4;
  constexpr size_t smem = (size_t)BULK_NW * 2 * 32 * D * 4 + BULK_NW * 2 * 8;
  static bool configured = false;
  if (!configured) {
    GODE_CHECK_CUDA(cudaFuncSetAttribute(k_spmm_bulk<VPL, false, ASYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GODE_CHECK_CUDA(cudaFuncSetAttribute(k_spmm_bulk<VPL, true, ASYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  constexpr int UPB = BULK_NW * BULK_R;  // units per block
  if (A.n_rows > 0) {
    unsigned grid = static_cast<unsigned>((A.n_rows + UPB - 1) / UPB);
    k_spmm_bulk<VPL, false, ASYNC><<<grid, BULK_NW * 32, smem, st>>>(A.n_rows, A.rowptr, A.colidx, A.vals, nullptr, nullptr, 0, X, ldx,
                                                              Y, ldy, nullptr, ep);
    GODE_LAUNCH_CHECK();
  }
  if (A.n_heavy > 0) {
    unsigned g1 = static_cast<unsigned>((A.n_chunks + UPB - 1) / UPB);
    k_spmm_bulk<VPL, true, ASYNC><<<g1, BULK_NW * 32, smem, st>>>(A.n_chunks, A.rowptr, A.colidx, A.vals, A.heavy_rows,
                                                           A.heavy_chunk_ptr, A.n_heavy, X, ldx, nullptr, 0, ws, ep);
    GODE_LAUNCH_CHECK();
    constexpr int RPB = 8;
    unsigned g2 = static_cast<unsigned>((A.n_heavy + RPB - 1) / RPB);
    k_spmm_heavy_finish<32, VPL><<<g2, 256, 0, st>>>(A.n_heavy, A.heavy_rows, A.heavy_chunk_ptr, ws, Y, ldy, ep);
    GODE_LAUNCH_CHECK();
  }
  return GODE_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static bool vec_width(int d) { return d == 8 || d == 16 || d == 32 || d == 64 || d == 128 || d == 256; }

size_t spmm_ws_bytes(const gode_csr_t& A, int d) {
  if (A.n_heavy <= 0 || !vec_width(d)) return 0;
  return align_up(sizeof(float) * static_cast<size_t>(A.n_chunks) * d, 256);
}

int spmm_dispatch(const gode_csr_t& A, const float* X, int64_t ldx, int32_t d, float* Y, int64_t ldy,
                  const gode_spmm_epilogue_t& ep, void* ws, size_t ws_bytes, cudaStream_t st, int64_t row_begin, int64_t row_end) {
  bool vec_ok = vec_width(d) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(X) && aligned16(Y) && aligned16(ep.bias) &&
                aligned16(ep.residual) && aligned16(ep.y0) && aligned16(ep.ynext) && aligned16(ep.mask_src) &&
                aligned16(ep.gp_out) && aligned16(ep.acc_in) && aligned16(ep.second.out) && aligned16(ws);
  for (int j = 0; j < ep.n_prev; ++j) vec_ok = vec_ok && aligned16(ep.kprev[j]);
  if (vec_ok && A.n_heavy > 0 && (!ws || ws_bytes < spmm_ws_bytes(A, d) || !A.heavy_rows || !A.heavy_chunk_ptr)) {
    set_error("spmm: heavy-row workspace missing or too small (%zu < %zu)", ws_bytes, spmm_ws_bytes(A, d));
    return GODE_EWORKSPACE;
  }
  const bool sub = row_begin != 0 || (row_end >= 0 && row_end != A.n_rows);
  if (sub && !(vec_ok && d == 128 && ldx == 128)) {
    set_error("spmm: a row-range call needs d = 128 with contiguous, 16-byte aligned rows (the k_spmm_t2 path)");
    return GODE_EINVAL;
  }
  if (vec_ok) {
    float* w = static_cast<float*>(ws);
    static const int use_bulk = [] {
      const char* e = getenv("GODE_SPMM_BULK");
      return e ? atoi(e) : 0;   // 0: register-staged (default), 1: LDGSTS pipeline, 2: TMA bulk-copy pipeline
    }();
    if (use_bulk == 1 && d == 128 && !sub) return launch_bulk<1, 1>(A, X, ldx, Y, ldy, ep, w, st);
    if (use_bulk == 2 && d == 128 && !sub) return launch_bulk<1, 0>(A, X, ldx, Y, ldy, ep, w, st);
    switch (d) {
      case 8: return launch_vec<2, 1>(A, X, ldx, Y, ldy, ep, w, st);
      case 16: return launch_vec<4, 1>(A, X, ldx, Y, ldy, ep, w, st);
      case 32: return launch_vec<8, 1>(A, X, ldx, Y, ldy, ep, w, st);
      case 64: return launch_vec<16, 1>(A, X, ldx, Y, ldy, ep, w, st);
      case 128: return launch_vec<32, 1>(A, X, ldx, Y, ldy, ep, w, st, row_begin, row_end);
      case 256: return launch_vec<32, 2>(A, X, ldx, Y, ldy, ep, w, st);
      default: break;
    }
  }
  GODE_REQUIRE(!ep.push.ptr, "spmm: a fused halo push needs a vectorised width (8..256, 16-byte aligned operands)");
  GODE_REQUIRE(row_begin == 0 && (row_end < 0 || row_end == A.n_rows), "spmm: a row-range call needs the vectorised d = 128 path");
  if (A.n_rows > 0) {
    unsigned grid = static_cast<unsigned>((A.n_rows + 7) / 8);
    k_spmm_generic<<<grid, 256, 0, st>>>(A.n_rows, A.rowptr, A.colidx, A.vals, X, ldx, d, Y, ldy, ep);
    GODE_LAUNCH_CHECK();
  }
  return GODE_OK;
}

}  // namespace gode

extern "C" size_t gode_spmm_workspace_bytes(const gode_csr_t* A, int32_t d) {
  if (!A) return 0;
  return gode::spmm_ws_bytes(*A, d);
}

extern "C" int gode_spmm_csr_f32(const gode_csr_t* A, const float* X, int64_t ldx, int32_t d, float* Y, int64_t ldy,
                                 const gode_spmm_epilogue_t* epi, void* ws, size_t ws_bytes, void* stream) {
  using namespace gode;
  GODE_REQUIRE(A && A->n_rows >= 0 && d > 0 && ldx >= d && (Y == nullptr || ldy >= d), "spmm: bad shape");
  GODE_REQUIRE(A->rowptr && X, "spmm: null pointer");
  gode_spmm_epilogue_t ep;
  if (epi) {
    ep = *epi;
  } else {
    memset(&ep, 0, sizeof(ep));
  }
  GODE_REQUIRE(ep.n_prev >= 0 && ep.n_prev <= GODE_MAX_STAGES, "spmm: n_prev out of range");
  GODE_REQUIRE((!ep.ynext && !ep.second.out) || ep.y0, "spmm: ynext / second.out need y0");
  GODE_REQUIRE(!ep.gp_out || ep.mask_src, "spmm: gp_out needs mask_src");
  GODE_REQUIRE(Y || ep.ynext || ep.gp_out || ep.second.out, "spmm: no output requested");
  if (!Y && ldy < d) ldy = d;  // every epilogue operand shares the leading dimension ldy
  ProfScope prof(GODE_PROF_OTHER, as_stream(stream));
  return spmm_dispatch(*A, X, ldx, d, Y, ldy, ep, ws, ws_bytes, as_stream(stream));
}
