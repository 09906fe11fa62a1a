// libgode: tensor-core (tcgen05 / TMEM) kernels for the d x d products of the GCN ODE function.
//
// Reference call sites: torch.mm(input, weight) inside FixedGraphConvolution.forward (GCN/layers.py:70) applied to
// [t || GroupNorm(x)] (GCN/models.py:175-178), and the two products its autograd issues (dS W^T and z^T dS).
//
// fp32 result on the 5th-generation tensor cores: every fp32 operand is split x = hi + lo with hi = tf32(x)
// (cvt.rna) and lo = tf32(x - hi) (rounded too: the MMA truncates its operands), and three kind::tf32 MMAs compute
// hi*hi + lo*hi + hi*lo (the dropped lo*lo term is 2^-22 relative).  The correction terms go to their own TMEM
// accumulator and the epilogue adds it to the hi*hi one with round-to-nearest fp32 adds (the MMA's own accumulate loses
// low bits at every K step); GODE_TC_ACC=19 additionally splits the hi*hi products over two accumulators, which makes the
// result as accurate as cuBLAS SGEMM (table at GODE_TC_ACC_DEFAULT below).  GODE_PREC_TF32 issues the hi*hi pass only.
//
// k_rows_tc  (M = 128 rows per tile, N = K = D):     Out[r,:] = A[r,:] * B + rowvec
//     MODE 0  transform : A = xhat(y) (GroupNorm statistics computed on load, affine folded into B),
//                         B = diag(gamma) W[1:,:], rowvec = t W[0,:] + beta^T W[1:,:]            -> S
//     MODE 1  input grad: A = dS, B = W[1:,:]^T                                                    -> dz
//   One persistent CTA per SM.  B (hi and lo) stays resident in shared memory in the UMMA canonical
//   K-major no-swizzle layout; the A tile is staged half of K at a time (registers -> shared, converted and
//   split on the way) while the next half is already in flight from HBM; one elected thread issues the MMAs;
//   the accumulator is read back with tcgen05.ld, staged through shared memory and stored as full 512 B rows.
//
// k_wgrad_tc (M = N = D channels, K = rows):         P = xhat(y)^T * dS   (accumulated over all rows of a CTA)
//   Both operands are row-major [rows, D] tiles used as MN-major UMMA operands -- the same physical layout.
//
// Shared-memory layout of an [R rows, C channels] fp32 tile (bytes):
//     off(r, c) = (r / 8) * (C * 32) + (c / 4) * 128 + (r % 8) * 16 + (c % 4) * 4
//   = 8-row x 16-byte core matrices; as a K-major operand (rows = M/N, channels = K): LBO = 128, SBO = C*32;
//   as an MN-major operand (channels = M/N, rows = K): SBO = 128, LBO = C*32.
#include "internal.cuh"
#include <stdlib.h>
#include <string.h>

namespace gode {

namespace tc {

constexpr int THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
  // | base_offset=0 | lbo_mode=0 | layout_type=SWIZZLE_NONE(0) [61,64)
  return static_cast<uint64_t>((addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo >> 4) << 16) |
         (static_cast<uint64_t>(sbo >> 4) << 32) | (1ull << 46);
}

// K-major operand in the 128-byte-swizzled canonical layout (cute::UMMA Layout_K_SW128_Atom): a row is 128 contiguous
// bytes (32 tf32), 16-byte chunk c of row r sits at chunk position c ^ (r % 8), 8-row groups are SBO = 1024 bytes
// apart, LBO = 16 bytes; K steps advance the start address by raw bytes inside the row.  The tile must be 1024-byte
// aligned.  layout_type = SWIZZLE_128B (2) in bits [61,64).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr) {
  return static_cast<uint64_t>((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}

__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
  // InstrDescriptor: c_format=F32 (1<<4) | a_format=TF32 (2<<7) | b_format=TF32 (2<<10) | a_major<<15 | b_major<<16
  // | (N>>3)<<17 | (M>>4)<<24
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TC_DONE;\n"
      "bra TC_WAIT;\n"
      "TC_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread t of warp w reads TMEM lane 32*(w%4)+t
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Streams that are read once are latency-bound when only the registers of one staged tile are in flight per SM
// (32 KB / ~1.5 us = 3 TB/s over 148 SMs -- what these kernels measured).  Asking L2 for the lines of later tiles costs
// no registers and turns the eventual loads into L2 hits.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// x = hi + lo with hi = tf32(x) (round to nearest).  lo = x - hi is exact in fp32 but carries up to 13 significant bits;
// the MMA reads only its top 11 (it TRUNCATES the operand to tf32), a biased error of up to 2^-23 |x|.  With `rnd` the
// residual is itself rounded to nearest tf32 (as CUTLASS's 3xTF32 does), halving that error and removing the bias.
__device__ __forceinline__ void split4(const float4& x, float4& hi, float4& lo, bool rnd = true) {
  hi.x = tf32_hi(x.x); hi.y = tf32_hi(x.y); hi.z = tf32_hi(x.z); hi.w = tf32_hi(x.w);
  lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
  if (rnd) { lo.x = tf32_hi(lo.x); lo.y = tf32_hi(lo.y); lo.z = tf32_hi(lo.z); lo.w = tf32_hi(lo.w); }
}

// GroupNorm statistics of one 16-byte chunk (4 channels): CPG = 4 -> one group, CPG = 2 -> two groups; no affine
template <int CPG>
__device__ __forceinline__ float4 normalize4(float4 x, float eps) {
  if (CPG == 4) {
    const float mean = (x.x + x.y + x.z + x.w) * 0.25f;
    const float a = x.x - mean, b = x.y - mean, c = x.z - mean, d = x.w - mean;
    const float rstd = 1.0f / sqrtf((a * a + b * b + c * c + d * d) * 0.25f + eps);
    return make_float4(a * rstd, b * rstd, c * rstd, d * rstd);
  } else if (CPG == 2) {
    const float m0 = (x.x + x.y) * 0.5f, m1 = (x.z + x.w) * 0.5f;
    const float a = x.x - m0, b = x.y - m0, c = x.z - m1, d = x.w - m1;
    const float r0 = 1.0f / sqrtf((a * a + b * b) * 0.5f + eps), r1 = 1.0f / sqrtf((c * c + d * d) * 0.5f + eps);
    return make_float4(a * r0, b * r0, c * r1, d * r1);
  }
  return x;
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------
// Out[r, :] = A[r, :] * B + rowvec,  rows tiled by 128
// ------------------------------------------------------------------------------------------------
template <int D, int CPG, int MODE>
__global__ void __launch_bounds__(tc::THREADS, 1)
k_rows_tc(int64_t n_rows, const float* __restrict__ X, float* __restrict__ Out, const float* __restrict__ W /*[D+1, D]*/,
          const float* __restrict__ gamma, const float* __restrict__ beta, float t, float eps, int passes,
          const gode_push_route_t push /*MODE 0: rows peers reference are also stored into their halo tails*/) {
  using namespace tc;
  constexpr int KH = D / 2;                 // channels per staged half of K
  constexpr int NI = (16 * (KH / 16)) / 8;  // warp-instructions (8 rows x 4 chunks) per warp per half tile
  constexpr uint32_t RSB = D * 32;          // 8-row-group stride of the B tile (full K)
  constexpr uint32_t RSA = KH * 32;         // 8-row-group stride of an A half tile
  constexpr uint32_t B_BYTES = D * D * 4;
  constexpr uint32_t A_BYTES = 128 * KH * 4;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sBhi = smem;
  unsigned char* sBlo = smem + B_BYTES;
  unsigned char* sAhi = smem + 2 * B_BYTES;
  unsigned char* sAlo = sAhi + A_BYTES;               // sAhi..sAlo+A_BYTES doubles as the epilogue staging buffer
  float* sRow = reinterpret_cast<float*>(sAlo + A_BYTES);   // [D] rowvec
  uint64_t* bar = reinterpret_cast<uint64_t*>(sRow + D);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_tiles = (n_rows + 127) / 128;
  if ((int64_t)blockIdx.x >= n_tiles) return;

  // ---- one-time setup: barrier, TMEM, B operand, row vector -----------------------------------
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 2 * D);   // columns [0, D): hi*hi; [D, 2D): the correction terms (added in the epilogue)
  const float* W1 = W + D;   // rows 1..D of the [D+1, D] weight
  for (int n = tid; n < D; n += THREADS) {
    float r = 0.f;
    if (MODE == 0) {
      r = t * __ldg(W + n);
      for (int k = 0; k < D; ++k) r = fmaf(__ldg(beta + k), __ldg(W1 + (int64_t)k * D + n), r);
    }
    sRow[n] = r;
  }
  // B[n][k]: MODE 0 -> gamma[k] * W1[k][n] ; MODE 1 -> W1[n][k].  lane -> (n%8, 4 k-chunks); conflict-free stores.
  for (int g = warp; g < (D / 8) * (D / 16); g += THREADS / 32) {
    const int ng = g / (D / 16), cg = g % (D / 16);
    const int n = ng * 8 + (lane & 7), kc = cg * 4 + (lane >> 3);
    float4 b;
    if (MODE == 0) {
      const int k = kc * 4;
      b.x = __ldg(gamma + k) * __ldg(W1 + (int64_t)k * D + n);
      b.y = __ldg(gamma + k + 1) * __ldg(W1 + (int64_t)(k + 1) * D + n);
      b.z = __ldg(gamma + k + 2) * __ldg(W1 + (int64_t)(k + 2) * D + n);
      b.w = __ldg(gamma + k + 3) * __ldg(W1 + (int64_t)(k + 3) * D + n);
    } else {
      b = __ldg(reinterpret_cast<const float4*>(W1 + (int64_t)n * D + kc * 4));
    }
    float4 hi, lo;
    split4(b, hi, lo);
    const uint32_t off = ng * RSB + kc * 128 + (lane & 7) * 16;
    *reinterpret_cast<float4*>(sBhi + off) = hi;
    *reinterpret_cast<float4*>(sBlo + off) = lo;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  constexpr uint32_t IDESC = make_idesc(128, D, false, false);
  const uint32_t aHi = smem_u32(sAhi), aLo = smem_u32(sAlo), bHi = smem_u32(sBhi), bLo = smem_u32(sBlo);

  // ---- software pipeline over (tile, half) stages ------------------------------------------------
  float4 raw[NI];
  auto load_raw = [&](int64_t tile, int half) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int g = warp * NI + i;
      const int rg = g / (KH / 16), cg = g % (KH / 16);
      const int64_t row = tile * 128 + rg * 8 + (lane & 7);
      const int kc = cg * 4 + (lane >> 3);
      raw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < n_rows) raw[i] = __ldcs(reinterpret_cast<const float4*>(X + row * D + half * KH + kc * 4));
    }
  };
  auto store_half = [&]() {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int g = warp * NI + i;
      const int rg = g / (KH / 16), cg = g % (KH / 16);
      const int kc = cg * 4 + (lane >> 3);
      float4 hi, lo;
      split4(normalize4<CPG>(raw[i], eps), hi, lo);
      const uint32_t off = rg * RSA + kc * 128 + (lane & 7) * 16;
      *reinterpret_cast<float4*>(sAhi + off) = hi;
      *reinterpret_cast<float4*>(sAlo + off) = lo;
    }
  };

  uint32_t parity = 0;
  int64_t tile = blockIdx.x;
  load_raw(tile, 0);
  bool pending = false;   // an MMA batch has been committed and not yet waited for
  for (; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      if (pending) {   // the A buffer is still being read by the previous half's MMAs
        mbar_wait(bar, parity);
        parity ^= 1u;
        pending = false;
      }
      store_half();
      // next stage's loads fly while this half is multiplied
      if (half == 0) load_raw(tile, 1);
      else if (tile + gridDim.x < n_tiles) load_raw(tile + gridDim.x, 0);
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < KH / 8; ++s) {
          const uint32_t ks = half * (KH / 8) + s;
          const uint64_t da_hi = make_desc(aHi + s * 256, 128, RSA), da_lo = make_desc(aLo + s * 256, 128, RSA);
          const uint64_t db_hi = make_desc(bHi + ks * 256, 128, RSB), db_lo = make_desc(bLo + ks * 256, 128, RSB);
          mma_tf32(tmem_d, da_hi, db_hi, IDESC, (half | s) != 0 ? 1u : 0u);
          if (passes == 3) {
            mma_tf32(tmem_d + D, da_lo, db_hi, IDESC, (half | s) != 0 ? 1u : 0u);
            mma_tf32(tmem_d + D, da_hi, db_lo, IDESC, 1u);
          }
        }
        mma_commit(bar);
      }
      pending = true;
    }
    // ---- epilogue: TMEM -> registers -> (+rowvec) -> staging smem -> 512 B row stores ------------
    mbar_wait(bar, parity);
    parity ^= 1u;
    pending = false;
    tc_fence_after();
    {
      constexpr int CH = D / 4;            // 16-byte chunks per output row
      constexpr int HC = D / 2;            // columns handled by one warp (two warps share a TMEM lane quarter)
      const int q = warp & 3, hc = warp >> 2;
      const int row = q * 32 + lane;
      float* stage = reinterpret_cast<float*>(sAhi);
#pragma unroll
      for (int cb = 0; cb < HC; cb += 32) {
        float v[32];
        tmem_ld32(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + hc * HC + cb, v);
        if (passes == 3) {
          float w[32];
          tmem_ld32(tmem_d + D + (static_cast<uint32_t>(q * 32) << 16) + hc * HC + cb, w);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += w[j];
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int col = hc * HC + cb + j;
          const int chunk = (col >> 2) ^ (row & (CH - 1));
          float4 o = make_float4(v[j] + sRow[col], v[j + 1] + sRow[col + 1], v[j + 2] + sRow[col + 2], v[j + 3] + sRow[col + 3]);
          *reinterpret_cast<float4*>(stage + row * D + chunk * 4) = o;
        }
      }
      tc_fence_before();
      __syncthreads();
      // each warp writes whole rows: lane -> 16-byte chunk(s) of the row
      for (int r = warp; r < 128; r += THREADS / 32) {
        const int64_t grow = tile * 128 + r;
        if (grow >= n_rows) break;
        int p0 = 0, p1 = 0;
        if (MODE == 0 && push.ptr) {
          p0 = __ldg(push.ptr + grow);
          p1 = __ldg(push.ptr + grow + 1);
        }
#pragma unroll
        for (int c = lane; c < CH; c += 32) {
          const float4 o = *reinterpret_cast<const float4*>(stage + r * D + ((c ^ (r & (CH - 1))) * 4));
          __stcs(reinterpret_cast<float4*>(Out + grow * D + c * 4), o);
          for (int e = p0; e < p1; ++e) {   // fused halo push over NVLink (posted stores)
            const int64_t ent = __ldg(push.ent + e);
            *reinterpret_cast<float4*>(push.base[ent >> 40] + (ent & 0xFFFFFFFFFFll) * D + c * 4) = o;
          }
        }
      }
      __syncthreads();   // staging buffer is the A buffer of the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 2 * D);
}

// ------------------------------------------------------------------------------------------------
// k_rows_ws: the same product as k_rows_tc, warp-specialised so that the three phases of a tile overlap:
//   warps 0-7 (producers): HBM -> registers (two quarter-K stages ahead) -> normalise / split -> shared memory,
//                          then arrive on the stage's "full" mbarrier (ncu: four producer warps were issue-bound
//                          on the conversion and everything else waited for them);
//   warp 12 (one thread) : waits for a full stage, issues its K/8 steps x 3 passes of tcgen05.mma into one of TWO
//                          TMEM accumulators, commits the stage's "empty" barrier;
//   warps 8-11 (epilogue): previous tile's accumulator -> registers -> (+rowvec) -> staging -> full-sector row stores
//                          (and, MODE 0 with a push route, the peers' halo tails over NVLink).
// A stage is released by the tcgen05.commit of the MMAs that read it, an accumulator by the epilogue warps once
// they have read it out, so the tensor pipe, the shared-memory stores and the HBM streams of consecutive tiles run
// concurrently instead of back to back as in k_rows_tc.
// Shared memory: B hi|lo (2 x D*D*4) + 2 A stages x (hi|lo) of 128 rows x 32 channels (64 KB) + 32 KB staging.
// ------------------------------------------------------------------------------------------------
namespace tc {
__device__ __forceinline__ void bar_sync_named(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
}  // namespace tc

#ifndef GODE_TC_ACC_DEFAULT
// Accuracy configurations of k_rows_ws, measured against fp64 on a B200 (profiles/r02_transform_accuracy.jsonl, N = 262 144,
// d = 128; error / max|S| as rms, max; ReLU masks of 33.5 M that differ from the fp64 result; kernel ms at N = 10 M):
//      0  round 1: one accumulator, truncated residuals   1.72e-7  1.29e-6   4 flips   2.4 ms   (gradients off by 2e-4..6e-4)
//      1  + residuals rounded to tf32                     1.71e-7  1.26e-6   6 flips   2.4 ms   (gradients 7e-7)
//      3  + correction terms in their own accumulator     6.2e-8   6.1e-7    2 flips   2.5 ms   <- input gradient (MODE 1)
//     19  + two hi*hi accumulators (single-buffered)      3.9e-8   4.7e-7    0 flips   3.2 ms   <- transform (MODE 0)
//         cuBLAS SGEMM (torch.mm) / libgode SIMT FFMA      3.9e-8   5.0e-7    - / 1     - / 15 ms
// The transform decides the ReLU masks, so it runs the configuration that matches the fp32 library GEMM digit for digit
// (with 3, one element of the d = 128 rk4 reference fixture sits at 1.07e-5, just outside the 1e-5 bar); the input
// gradient decides no mask and keeps the double-buffered 3.  GODE_TC_ACC overrides both.
#define GODE_TC_ACC_DEFAULT 19
#endif

namespace tc {
constexpr int WS_PRODUCERS = 256;               // warps 0-7
constexpr int WS_THREADS = WS_PRODUCERS + 128 + 32;   // + epilogue warps 8-11 + MMA warp 12

// GroupNorm of one 16-byte chunk with rsqrtf (MUFU.RSQ, <= 2 ulp): the producers are issue-bound, and
// 1 / sqrtf costs ~15 instructions per chunk more for digits the 1e-5 bar does not see
template <int CPG>
__device__ __forceinline__ float4 normalize4_fast(float4 x, float eps) {
  if (CPG == 4) {
    const float mean = (x.x + x.y + x.z + x.w) * 0.25f;
    const float a = x.x - mean, b = x.y - mean, c = x.z - mean, d = x.w - mean;
    const float rstd = rsqrtf((a * a + b * b + c * c + d * d) * 0.25f + eps);
    return make_float4(a * rstd, b * rstd, c * rstd, d * rstd);
  }
  return normalize4<CPG>(x, eps);
}
}  // namespace tc

template <int D, int CPG, int MODE, int CFG /*>= 0: compile-time accuracy configuration; -1: the run-time argument*/>
__global__ void __launch_bounds__(tc::WS_THREADS, 1)
k_rows_ws(int64_t n_rows, const float* __restrict__ X, float* __restrict__ Out, const float* __restrict__ W /*[D+1, D]*/,
          const float* __restrict__ gamma, const float* __restrict__ beta, float t, float eps, int passes,
          const gode_push_route_t push, const int pf_tiles /*L2 prefetch distance in tiles of this CTA (0: off)*/,
          const int cfg_rt /*accuracy configuration of the 3xTF32 product, see rows_acc_cfg()*/,
          const int64_t ldx /*row stride of X*/, const int64_t ldo /*row stride of Out*/, const int64_t ldw /*row stride of W*/) {
  using namespace tc;
  static_assert(D == 128, "k_rows_ws is laid out for 128 channels");
  // cfg: bit 0 round the lo residuals to tf32 | bit 1 correction terms (lo*hi, hi*lo[, lo*lo]) in their own TMEM
  // accumulator, added to the hi*hi one by the epilogue with round-to-nearest fp32 adds | bit 2 also issue lo*lo |
  // bit 3 1/sqrtf instead of rsqrtf in the GroupNorm | bits 4-5: (number of hi*hi accumulators over K) - 1; more than
  // one leaves no TMEM for double buffering (512 columns = 4 accumulators of 128)
  // (the single MMA-issuing thread evaluates these per instruction: with a run-time configuration its address arithmetic
  // became the tile's critical path -- +27 % kernel time -- so the shipped configurations are template constants)
  const int cfg = CFG >= 0 ? CFG : cfg_rt;
  const bool rnd_lo = cfg & 1, sep = (cfg & 2) != 0, lolo = (cfg & 4) != 0, precise_gn = (cfg & 8) != 0;
  const int nmain = 1 + ((cfg >> 4) & 3);
  const int nacc = nmain + (sep ? 1 : 0);
  const int nbuf = nacc <= 2 ? 2 : 1;
  constexpr int KQ = 32;                        // channels per A stage (a quarter of K)
  constexpr int NQ = D / KQ;                    // stages per tile
  constexpr int NI = 4;                         // warp-instructions (8 rows x 4 chunks) per producer warp per stage
  constexpr uint32_t RSB = D * 32;              // 8-row-group stride of the B tile (full K)
  constexpr uint32_t RSA = KQ * 32;             // 8-row-group stride of an A stage (8 rows x 128 bytes = one swizzle atom)
  constexpr uint32_t B_BYTES = D * D * 4;
  constexpr uint32_t A_BYTES = 128 * KQ * 4;    // one of hi / lo of one stage
  constexpr uint32_t STG_BYTES = 128 * (D / 2) * 4;   // staging: 128 rows x half of the columns
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sBhi = smem;
  unsigned char* sBlo = smem + B_BYTES;
  unsigned char* sA = smem + 2 * B_BYTES;                 // stage s: hi at sA + s*2*A_BYTES, lo right behind it
  float* stage = reinterpret_cast<float*>(sA + 4 * A_BYTES);
  float* sRow = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(stage) + STG_BYTES);   // [D] rowvec
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRow + D);   // [0,1] a_full, [2,3] a_empty, [4,5] acc_full, [6,7] acc_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_tiles = (n_rows + 127) / 128;
  if ((int64_t)blockIdx.x >= n_tiles) return;
  const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  // ---- one-time setup (all warps): barriers, TMEM (two accumulators), B operand, row vector ------------------
  if (tid == 0) {
    mbar_init(&bars[0], WS_PRODUCERS);
    mbar_init(&bars[1], WS_PRODUCERS);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], 1);
    mbar_init(&bars[4], 1);
    mbar_init(&bars[5], 1);
    mbar_init(&bars[6], 128);
    mbar_init(&bars[7], 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 4 * D);
  const float* W1 = W + ldw;
  for (int n = tid; n < D; n += WS_THREADS) {
    float r = 0.f;
    if (MODE == 0) {
      r = t * __ldg(W + n);
      for (int k = 0; k < D; ++k) r = fmaf(__ldg(beta + k), __ldg(W1 + (int64_t)k * ldw + n), r);
    }
    sRow[n] = r;
  }
  for (int g = warp; g < (D / 8) * (D / 16); g += WS_THREADS / 32) {
    const int ng = g / (D / 16), cg = g % (D / 16);
    const int n = ng * 8 + (lane & 7), kc = cg * 4 + (lane >> 3);
    float4 b;
    if (MODE == 0) {
      const int k = kc * 4;
      b.x = __ldg(gamma + k) * __ldg(W1 + (int64_t)k * ldw + n);
      b.y = __ldg(gamma + k + 1) * __ldg(W1 + (int64_t)(k + 1) * ldw + n);
      b.z = __ldg(gamma + k + 2) * __ldg(W1 + (int64_t)(k + 2) * ldw + n);
      b.w = __ldg(gamma + k + 3) * __ldg(W1 + (int64_t)(k + 3) * ldw + n);
    } else {
      b = __ldg(reinterpret_cast<const float4*>(W1 + (int64_t)n * ldw + kc * 4));
    }
    float4 hi, lo;
    split4(b, hi, lo, rnd_lo);
    const uint32_t off = ng * RSB + kc * 128 + (lane & 7) * 16;
    *reinterpret_cast<float4*>(sBhi + off) = hi;
    *reinterpret_cast<float4*>(sBlo + off) = lo;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t IDESC = make_idesc(128, D, false, false);
  const int64_t n_steps = my_tiles * NQ;

  if (warp < 8) {
    // =============================== producers ============================================================
    float4 raw[2][NI];
    // A warp instruction covers 4 rows x 128 bytes: the 8 lanes of a quarter warp read the 8 chunks of ONE row, so every
    // quarter-warp request is one 128-byte line (with lane -> row, as in k_rows_tc, each quarter warp touched 8 lines
    // and the L1 LSU pipe -- 80 % busy in ncu -- spent 32 wavefronts per LDG.128 instead of 4).
    auto load_raw = [&](float4 (&r)[NI], int64_t step) {
      const int64_t tile = blockIdx.x + (step / NQ) * (int64_t)gridDim.x;
      const int q = static_cast<int>(step % NQ);
      const bool live = step < n_steps;
      const bool pf = pf_tiles > 0 && step + pf_tiles * NQ < n_steps && (lane & 7) == 0;   // one request per row line
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int64_t row = tile * 128 + (warp * NI + i) * 4 + (lane >> 3);
        const float* src = X + row * ldx + q * KQ + (lane & 7) * 4;
        r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live && row < n_rows) r[i] = __ldcs(reinterpret_cast<const float4*>(src));
        if (pf) {                                     // same bytes of the tile pf_tiles visits ahead
          const int64_t prow = row + pf_tiles * (int64_t)gridDim.x * 128;
          if (prow < n_rows) prefetch_l2(src + pf_tiles * (int64_t)gridDim.x * 128 * ldx);
        }
      }
    };
    auto do_step = [&](float4 (&r)[NI], int64_t step) {
      const int s = static_cast<int>(step & 1);
      const int64_t use = step >> 1;               // how many times stage s has been filled before
      if (use >= 1) mbar_wait(&bars[2 + s], static_cast<uint32_t>((use - 1) & 1));   // MMAs of its previous contents retired
      unsigned char* aHi = sA + (size_t)s * 2 * A_BYTES;
      unsigned char* aLo = aHi + A_BYTES;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int rr = (warp * NI + i) * 4 + (lane >> 3);      // row of the tile
        float4 hi, lo;
        split4(precise_gn ? normalize4<CPG>(r[i], eps) : normalize4_fast<CPG>(r[i], eps), hi, lo, rnd_lo);
        // 128-byte-swizzled K-major stage: a quarter warp fills one 128-byte row -> conflict-free
        const uint32_t off = (rr >> 3) * RSA + (rr & 7) * 128 + (((lane & 7) ^ (rr & 7)) << 4);
        *reinterpret_cast<float4*>(aHi + off) = hi;
        *reinterpret_cast<float4*>(aLo + off) = lo;
      }
      load_raw(r, step + 2);                       // two stages ahead, into the registers just consumed
      fence_async_smem();
      mbar_arrive(&bars[s]);                       // a_full[s]: 256 arrivals
    };
    load_raw(raw[0], 0);
    load_raw(raw[1], 1);
    for (int64_t step = 0; step < n_steps; step += 2) {   // n_steps is a multiple of NQ = 4
      do_step(raw[0], step);
      do_step(raw[1], step + 1);
    }
  } else if (warp == 12) {
    // =============================== MMA issue (one thread) ===============================================
    if (lane == 0) {
      const uint32_t bHi = smem_u32(sBhi), bLo = smem_u32(sBlo);
      uint32_t started = 0;                          // accumulators of the current tile that hold a value already
      for (int64_t step = 0; step < n_steps; ++step) {
        const int s = static_cast<int>(step & 1);
        const int64_t it = step / NQ;
        const int q = static_cast<int>(step % NQ);
        const int b = static_cast<int>(it % nbuf);
        const int64_t use = it / nbuf;               // how many tiles have used accumulator set b before
        if (q == 0) {
          started = 0;
          if (use >= 1) mbar_wait(&bars[6 + b], static_cast<uint32_t>((use - 1) & 1));   // epilogue has read it out
        }
        mbar_wait(&bars[s], static_cast<uint32_t>((step >> 1) & 1));                                 // stage s is stored
        tc_fence_after();
        const uint32_t aHi = smem_u32(sA + (size_t)s * 2 * A_BYTES), aLo = aHi + A_BYTES;
        const uint32_t tmem_set = tmem_base + b * nacc * D;
#pragma unroll
        for (int ks = 0; ks < KQ / 8; ++ks) {
          const uint32_t kb = q * (KQ / 8) + ks;     // K step of the tile, 0..15
          const uint64_t da_hi = make_desc_sw128(aHi + ks * 32), da_lo = make_desc_sw128(aLo + ks * 32);
          const uint64_t db_hi = make_desc(bHi + kb * 256, 128, RSB), db_lo = make_desc(bLo + kb * 256, 128, RSB);
          const uint32_t am = (kb * nmain) >> 4;     // hi*hi accumulator of this K step
          mma_tf32(tmem_set + am * D, da_hi, db_hi, IDESC, (started >> am) & 1u);
          started |= 1u << am;
          if (passes == 3) {
            const uint32_t ac = sep ? static_cast<uint32_t>(nmain) : am;
            mma_tf32(tmem_set + ac * D, da_lo, db_hi, IDESC, (started >> ac) & 1u);
            started |= 1u << ac;
            mma_tf32(tmem_set + ac * D, da_hi, db_lo, IDESC, 1u);
            if (lolo) mma_tf32(tmem_set + ac * D, da_lo, db_lo, IDESC, 1u);
          }
        }
        mma_commit(&bars[2 + s]);                    // stage s is free once these MMAs have read it
        if (q == NQ - 1) mma_commit(&bars[4 + b]);   // accumulator set b is complete
      }
    }
  } else {
    // =============================== epilogue (warps 8-11) =================================================
    constexpr int HC = D / 2;                      // columns per staging pass
    constexpr int CHH = HC / 4;                    // 16-byte chunks per staged half row (16)
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int et = tid - WS_PRODUCERS;             // 0..127
    const int row = q * 32 + lane;                 // TMEM lane = row of the tile
    const bool do_push = MODE == 0 && push.ptr != nullptr;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t tile = blockIdx.x + it * (int64_t)gridDim.x;
      const int b = static_cast<int>(it % nbuf);
      mbar_wait(&bars[4 + b], static_cast<uint32_t>((it / nbuf) & 1));
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + b * nacc * D + (static_cast<uint32_t>(q * 32) << 16);
      const int nread = passes == 3 ? nacc : nmain;   // a single-pass product never touches the correction accumulator
      const bool full = tile * 128 + 128 <= n_rows;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int cb = 0; cb < HC; cb += 32) {
          float v[32];
          tmem_ld32(tmem_d + h * HC + cb, v);
          for (int a = 1; a < nread; ++a) {          // partial products over K and the correction terms: fp32 RN adds
            float w[32];
            tmem_ld32(tmem_d + a * D + h * HC + cb, w);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += w[j];
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int col = h * HC + cb + j;
            const int chunk = ((cb + j) >> 2) ^ (row & (CHH - 1));
            const float4 rv = *reinterpret_cast<const float4*>(sRow + col);
            *reinterpret_cast<float4*>(stage + row * HC + chunk * 4) = make_float4(v[j] + rv.x, v[j + 1] + rv.y, v[j + 2] + rv.z, v[j + 3] + rv.w);
          }
        }
        if (h == 1) {                              // the accumulator has been read out completely
          tc_fence_before();
          mbar_arrive(&bars[6 + b]);
        }
        bar_sync_named(2, 128);
        // 128 rows x 16 chunks: a warp instruction stores two 256-byte half rows; 16 chunks per thread
        float* obase = Out + (tile * 128) * ldo + h * HC;
        if (full && !do_push) {
#pragma unroll
          for (int i0 = 0; i0 < 16; i0 += 8) {
            float4 o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int idx = (i0 + i) * 128 + et;
              const int r = idx / CHH, c = idx % CHH;
              o[i] = *reinterpret_cast<const float4*>(stage + r * HC + ((c ^ (r & (CHH - 1))) * 4));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int idx = (i0 + i) * 128 + et;
              const int r = idx / CHH, c = idx % CHH;
              __stcs(reinterpret_cast<float4*>(obase + (int64_t)r * ldo + c * 4), o[i]);
            }
          }
        } else {
#pragma unroll 2
          for (int i = 0; i < 16; ++i) {
            const int idx = i * 128 + et;
            const int r = idx / CHH, c = idx % CHH;
            const int64_t grow = tile * 128 + r;
            if (grow < n_rows) {
              const float4 o = *reinterpret_cast<const float4*>(stage + r * HC + ((c ^ (r & (CHH - 1))) * 4));
              __stcs(reinterpret_cast<float4*>(obase + (int64_t)r * ldo + c * 4), o);
              if (do_push) {                       // fused halo push over NVLink (posted stores)
                const int p1 = __ldg(push.ptr + grow + 1);
                for (int e = __ldg(push.ptr + grow); e < p1; ++e) {
                  const int64_t ent = __ldg(push.ent + e);
                  *reinterpret_cast<float4*>(push.base[ent >> 40] + (ent & 0xFFFFFFFFFFll) * D + h * HC + c * 4) = o;
                }
              }
            }
          }
        }
        bar_sync_named(2, 128);                    // staging is reused by the next pass
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 4 * D);
}

// GODE_TC: bit 0 = transform, bit 1 = input gradient, bit 2 = weight gradient on tcgen05 (default all)
static int tc_mask() {   // read at every call (a getenv is ~100 ns): tools compare the paths inside one process
  const char* e = getenv("GODE_TC");
  return e ? atoi(e) : 7;
}
static bool tc_enabled() { return tc_mask() != 0; }

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ------------------------------------------------------------------------------------------------
// P[i, o] = sum_r xhat(y)[r, i] * dS[r, o]      (weight gradient before the GroupNorm affine is re-applied)
// ------------------------------------------------------------------------------------------------
// The reduction runs over ALL rows of a CTA (68 000 at N = 10 M: 8 400 K steps x 3 passes).  Left in one TMEM accumulator
// the MMA's own accumulate -- which drops low bits at every step -- cost 9e-5 relative L2 against fp64 at 2 M rows (round 2
// measurement).  So the accumulation is hierarchical: WG_DRAIN staged chunks (512 rows) go into one of two TMEM accumulator
// sets (hi*hi and the correction terms apart), and four drain warps add a finished set into fp32 running sums in shared
// memory with round-to-nearest adds while the MMAs continue into the other set.
namespace tc {
constexpr int WG_DRAIN = 16;                    // staged chunks (of 32 rows) per TMEM accumulation group
constexpr int WG_THREADS = THREADS + 32 + 128;  // producers (warps 0-7) + MMA issue (warp 8) + drain (warps 9-12)
}

template <int D, int CPG>
__global__ void __launch_bounds__(tc::WG_THREADS, 1)
k_wgrad_tc(int64_t n_rows, const float* __restrict__ Yin, const float* __restrict__ G, float* __restrict__ partial /*[grid][D][D]*/,
           float* __restrict__ cs_partial /*[grid][D]: column sums of G over this CTA's rows*/, float eps, int passes,
           const int pf_chunks /*L2 prefetch distance in chunks of this CTA (0: off)*/, const int rnd_lo /*round the lo residuals*/,
           const int64_t ldg /*row stride of G in floats*/, const int gcols /*valid columns of G (multiple of 4, <= D); the rest read 0*/) {
  using namespace tc;
  static_assert(D == 128, "the accumulator uses all 128 TMEM lanes");
  // Both operands are [rows, D] row-major in HBM but the reduction runs over rows, so they are transposed on the
  // way into shared memory and used as ordinary K-major operands: element (channel i, row r) of a staged tile at
  //     (i/8)*SBO + (r/4)*LBO + (i%8)*16 + (r%4)*4,   LBO = 144, SBO = 1168
  // (core matrices stay 128 B contiguous).  A warp instruction loads 4 rows x 128 bytes -- the 8 lanes of a quarter
  // warp read ONE 128-byte line (lane -> row made every quarter warp touch 8 lines: 32 LSU wavefronts per LDG.128
  // instead of 4, with the L1 LSU pipe 79 % busy in ncu) -- and its 4-byte transposing stores then go to channel
  // 4*(lane%8)+j, row lane/8: word bank = 16*(kc%2) + 4*(kc/2) + lane/8 + const with SBO = 16 bytes mod 128, i.e. 32
  // distinct banks.
  constexpr int RC = 32;                       // rows (= K) per staged chunk
  constexpr uint32_t LBO = 144, SBO = 1168;
  constexpr uint32_t MAT = (D / 8) * SBO;      // bytes of one staged operand tile
  constexpr int NI = ((RC / 8) * (D / 16)) / (THREADS / 32);   // warp-instructions per warp per operand
  extern __shared__ __align__(1024) unsigned char smem[];
  // stage b: [A_hi | A_lo | B_hi | B_lo] at smem + b * 4 * MAT
  // bars: [0],[1] stage free; [2] all MMAs done; [3],[4] stage full; [5],[6] accumulator set full; [7],[8] set drained
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * 4 * MAT);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float4* sums = reinterpret_cast<float4*>(smem + 2 * 4 * MAT + 128);   // running sums: [D / 4 column chunks][D channel rows]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_chunks = (n_rows + RC - 1) / RC;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], THREADS);
    mbar_init(&bars[4], THREADS);
    mbar_init(&bars[5], 1);
    mbar_init(&bars[6], 1);
    mbar_init(&bars[7], 128);
    mbar_init(&bars[8], 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 4 * D);   // two sets of (hi*hi | correction) accumulators
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  constexpr uint32_t IDESC = make_idesc(D, D, false, false);
  const int64_t my_chunks = (int64_t)blockIdx.x < n_chunks ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == THREADS / 32) {
    // ---- MMA issue (one thread of the ninth warp): a staged chunk is multiplied as soon as the eight producer warps
    // have arrived on its "full" barrier, so no producer ever waits for the issue (ncu: 16 % of the samples of the
    // single-barrier version sat in __syncthreads behind the issuing warp)
    if (lane == 0) {
      for (int64_t it = 0; it < my_chunks; ++it) {
        const int b = static_cast<int>(it & 1);
        const int64_t g = it / WG_DRAIN;                 // accumulation group
        const int ab = static_cast<int>(g & 1);          // its accumulator set
        const bool first = it % WG_DRAIN == 0;
        if (first && g >= 2) mbar_wait(&bars[7 + ab], static_cast<uint32_t>(((g >> 1) - 1) & 1));   // group g-2 has been drained
        mbar_wait(&bars[3 + b], static_cast<uint32_t>((it >> 1) & 1));
        tc_fence_after();
        const uint32_t base = smem_u32(smem + (size_t)b * 4 * MAT);
        const uint32_t t_main = tmem_d + ab * 2 * D, t_corr = t_main + D;
#pragma unroll
        for (int s = 0; s < RC / 8; ++s) {
          const uint32_t ko = s * 2 * LBO;   // 8 rows of K = two 16-byte chunks
          const uint64_t a_hi = make_desc(base + ko, LBO, SBO), a_lo = make_desc(base + MAT + ko, LBO, SBO);
          const uint64_t b_hi = make_desc(base + 2 * MAT + ko, LBO, SBO), b_lo = make_desc(base + 3 * MAT + ko, LBO, SBO);
          const uint32_t acc = (first && s == 0) ? 0u : 1u;
          mma_tf32(t_main, a_hi, b_hi, IDESC, acc);
          if (passes == 3) {
            mma_tf32(t_corr, a_lo, b_hi, IDESC, acc);
            mma_tf32(t_corr, a_hi, b_lo, IDESC, 1u);
          }
        }
        mma_commit(&bars[b]);
        if (it % WG_DRAIN == WG_DRAIN - 1 || it == my_chunks - 1) mma_commit(&bars[5 + ab]);   // this group's sums are complete
      }
      mma_commit(&bars[2]);   // arrives when every MMA issued above has completed
    }
    tc_fence_before();
    __syncthreads();
    return;
  }
  if (warp > THREADS / 32) {
    // ---- drain (warps 9-12): finished accumulator sets -> fp32 running sums in shared memory (RN adds), then -> partial
    const int q = warp & 3;                       // the TMEM lane quarter this warp may read
    const int row = q * 32 + lane;                // channel i = TMEM lane; a thread only ever touches its own row of sums
#pragma unroll 4
    for (int c4 = 0; c4 < D / 4; ++c4) sums[c4 * D + row] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t n_groups = (my_chunks + WG_DRAIN - 1) / WG_DRAIN;
    for (int64_t g = 0; g < n_groups; ++g) {
      const int ab = static_cast<int>(g & 1);
      mbar_wait(&bars[5 + ab], static_cast<uint32_t>((g >> 1) & 1));
      tc_fence_after();
      const uint32_t t_main = tmem_d + ab * 2 * D + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int cb = 0; cb < D; cb += 32) {
        float v[32];
        tmem_ld32(t_main + cb, v);
        if (passes == 3) {
          float w[32];
          tmem_ld32(t_main + D + cb, w);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += w[j];
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 t = sums[((cb + j) >> 2) * D + row];
          t.x += v[j]; t.y += v[j + 1]; t.z += v[j + 2]; t.w += v[j + 3];
          sums[((cb + j) >> 2) * D + row] = t;
        }
      }
      tc_fence_before();
      mbar_arrive(&bars[7 + ab]);
    }
    float* out = partial + (size_t)blockIdx.x * D * D + (size_t)row * D;
#pragma unroll 4
    for (int c4 = 0; c4 < D / 4; ++c4) *reinterpret_cast<float4*>(out + c4 * 4) = sums[c4 * D + row];
    tc_fence_before();
    __syncthreads();
    return;
  }

  float4 ra[NI], rb[NI], cs_acc[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) cs_acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load_raw = [&](int64_t chunk) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int64_t row = chunk * RC + warp * 4 + (lane >> 3);   // warp -> row quad, i -> 32-channel quarter
      const int kc = i * 8 + (lane & 7);
      ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      rb[i] = ra[i];
      if (row < n_rows) {
        ra[i] = __ldcs(reinterpret_cast<const float4*>(Yin + row * D + kc * 4));
        if (kc * 4 < gcols) rb[i] = __ldcs(reinterpret_cast<const float4*>(G + row * ldg + kc * 4));
      }
      const int64_t prow = row + pf_chunks * (int64_t)gridDim.x * RC;
      if (pf_chunks > 0 && prow < n_rows) {
        prefetch_l2(Yin + prow * D + kc * 4);
        if (kc * 4 < gcols) prefetch_l2(G + prow * ldg + kc * 4);
      }
    }
  };
  auto put = [&](unsigned char* tile, int r, int c0, const float4& v) {   // 4 channels of one row, transposed
    const uint32_t base = (r >> 2) * LBO + (r & 3) * 4;
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = c0 + j;
      *reinterpret_cast<float*>(tile + (i >> 3) * SBO + (i & 7) * 16 + base) = vv[j];
    }
  };
  auto store_stage = [&](int b, int64_t chunk) {
    unsigned char* base = smem + (size_t)b * 4 * MAT;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int kc = i * 8 + (lane & 7);
      const int r = warp * 4 + (lane >> 3);
      float4 xa = normalize4_fast<CPG>(ra[i], eps);
      if (chunk * RC + r >= n_rows) xa = make_float4(0.f, 0.f, 0.f, 0.f);   // xhat of a padding row is not zero by itself
      float4 hi, lo;
      split4(xa, hi, lo, rnd_lo != 0);
      put(base, r, kc * 4, hi);
      put(base + MAT, r, kc * 4, lo);
      cs_acc[i].x += rb[i].x; cs_acc[i].y += rb[i].y; cs_acc[i].z += rb[i].z; cs_acc[i].w += rb[i].w;   // padding rows are 0
      split4(rb[i], hi, lo, rnd_lo != 0);
      put(base + 2 * MAT, r, kc * 4, hi);
      put(base + 3 * MAT, r, kc * 4, lo);
    }
  };

  uint32_t parity[2] = {0u, 0u};
  int it = 0;
  int64_t chunk = blockIdx.x;
  if (chunk < n_chunks) load_raw(chunk);
  for (; chunk < n_chunks; chunk += gridDim.x, ++it) {
    const int b = it & 1;
    if (it >= 2) {   // the MMAs that read this stage two iterations ago must have retired
      mbar_wait(&bars[b], parity[b]);
      parity[b] ^= 1u;
    }
    store_stage(b, chunk);
    if (chunk + gridDim.x < n_chunks) load_raw(chunk + gridDim.x);
    fence_async_smem();
    mbar_arrive(&bars[3 + b]);
  }
  mbar_wait(&bars[2], 0);   // every MMA has retired: the operand stages are free for the column-sum reduction below
  tc_fence_after();
  // column sums of G over this CTA's rows: lanes sharing a 16-byte chunk differ in lane / 8 (the row inside a quad),
  // warps hold different row quads; fixed reduction order -> deterministic
  {
    float* scs = reinterpret_cast<float*>(smem);   // [8 warps][D]; the operand stages are free (all MMAs retired)
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      float4 v = cs_acc[i];
#pragma unroll
      for (int o = 8; o < 32; o <<= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
        v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
        v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
      }
      const int kc = i * 8 + (lane & 7);
      if ((lane >> 3) == 0) *reinterpret_cast<float4*>(scs + warp * D + kc * 4) = v;
    }
    bar_sync_named(1, THREADS);
    if (tid < D) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < THREADS / 32; ++w) t += scs[w * D + tid];
      cs_partial[(size_t)blockIdx.x * D + tid] = t;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 4 * D);
}

// cs[o] = sum_b cs_partial[b][o]   (column sums of dS; one block, fixed order)
__global__ void k_wgrad_cs(int nblk, int d, const float* __restrict__ cs_partial, float* __restrict__ cs) {
  const int o = threadIdx.x;
  if (o >= d) return;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += cs_partial[(size_t)b * d + o];
  cs[o] = s;
}

// gW1[i][o] = gamma[i] * sum_b partial[b][i][o] + beta[i] * cs[o]
__global__ void k_wgrad_finish(int nblk, int d, const float* __restrict__ partial, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ cs, float* __restrict__ gW1) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= d * d) return;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += partial[(size_t)b * d * d + idx];
  const int i = idx / d, o = idx % d;
  gW1[idx] = gamma[i] * s + beta[i] * cs[o];
}

bool wgrad_tc_supported(const gode_gcn_odefunc_t* f) { return (tc_mask() & 4) && f->d == 128 && f->groups == 32; }

size_t wgrad_tc_ws_bytes(const gode_gcn_odefunc_t* f) { return sizeof(float) * (size_t)sm_count() * (f->d + 1) * f->d; }

// gW1 = z^T dS with z = GroupNorm(y);  cs (output, [d]) = column sums of dS, accumulated in the same pass over dS
int wgrad_tc(const gode_gcn_odefunc_t* f, const float* y, const float* gS, float* cs, float* gW1, float* ws,
             size_t ws_bytes, cudaStream_t st) {
  constexpr int D = 128;
  GODE_REQUIRE(al16(y) && al16(gS) && al16(ws), "wgrad_tc: operands must be 16-byte aligned");
  const int64_t n_chunks = (f->A.n_rows + 31) / 32;
  int grid = static_cast<int>(n_chunks < persistent_ctas() ? n_chunks : persistent_ctas());
  if (grid < 1) grid = 1;
  if (ws_bytes < sizeof(float) * (size_t)grid * (D + 1) * D) {
    set_error("wgrad_tc: workspace too small");
    return GODE_EWORKSPACE;
  }
  float* cs_partial = ws + (size_t)grid * D * D;
  constexpr size_t smem = 2 * 4 * (size_t)(D / 8) * 1168 + 128 + (size_t)D * D * 4;
  static bool configured = false;
  if (!configured) {
    GODE_CHECK_CUDA(cudaFuncSetAttribute(k_wgrad_tc<D, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int passes = f->precision == GODE_PREC_TF32 ? 1 : 3;
  static const int pf = [] {
    const char* e = getenv("GODE_WGRAD_PREFETCH");   // L2 prefetch distance in units of 128 rows per CTA (default 2, 0 = off)
    return (e ? atoi(e) : 2) * 4;
  }();
  // GODE_WGRAD_RND=1 rounds the lo residuals to tf32 as the transform does.  Default off: this product decides no ReLU mask,
  // and its error against fp64 over 2 M rows is the same either way (1.232e-6 vs 1.234e-6) for 0.45 ms per launch.
  const char* rnd_env = getenv("GODE_WGRAD_RND");
  k_wgrad_tc<D, 4><<<grid, tc::WG_THREADS, smem, st>>>(f->A.n_rows, y, gS, ws, cs_partial, f->gn_eps, passes, pf,
                                                       rnd_env ? atoi(rnd_env) : 0, D, D);
  GODE_LAUNCH_CHECK();
  k_wgrad_cs<<<1, D, 0, st>>>(grid, D, cs_partial, cs);
  GODE_LAUNCH_CHECK();
  k_wgrad_finish<<<(D * D + 255) / 256, 256, 0, st>>>(grid, D, ws, f->gamma, f->beta, cs, gW1);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

// out[i][c0 + o] = sum_b partial[b][i][o],  cs[c0 + o] = sum_b cs_partial[b][o]      (o < gcols; fixed order)
__global__ void k_wgrad_finish_raw(int nblk, int d, int gcols, const float* __restrict__ partial, const float* __restrict__ cs_partial,
                                   float* __restrict__ out, int64_t ldo, float* __restrict__ cs) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (d + 1) * gcols) return;
  const int i = idx / gcols, o = idx % gcols;
  float s = 0.f;
  if (i < d) {
    for (int b = 0; b < nblk; ++b) s += partial[(size_t)b * d * d + (size_t)i * d + o];
    out[(size_t)i * ldo + o] = s;
  } else {
    for (int b = 0; b < nblk; ++b) s += cs_partial[(size_t)b * d + o];
    cs[o] = s;
  }
}

size_t gn_wgrad_ws_bytes(int d) { return sizeof(float) * (size_t)sm_count() * (d + 1) * d; }

// out[d, ncols] = xhat(y)^T * G,  cs[ncols] = column sums of G,  xhat = GroupNorm(y) before its affine (4 channels per
// group): the ODE-function weight gradient of a layer whose input is [t | GroupNorm(y)] and whose output is wider than d
// (the GAT projection, 2 * heads * oh + 2 * heads columns) -- k_wgrad_tc over 128-column blocks of G.
int gn_wgrad_tc(int64_t n, int d, int groups, float eps, const float* y, const float* G, int64_t ldg, int ncols, float* out,
                int64_t ldo, float* cs, float* ws, size_t ws_bytes, int precision, cudaStream_t st) {
  constexpr int D = 128;
  GODE_REQUIRE(d == D && groups == 32, "gn_wgrad: the tensor-core kernel covers d = 128 with 32 groups");
  GODE_REQUIRE(ncols > 0 && ncols % 4 == 0 && ldg % 4 == 0 && ldg >= ncols && ldo >= ncols, "gn_wgrad: ncols and ldg must be multiples of 4");
  GODE_REQUIRE(al16(y) && al16(G) && al16(ws), "gn_wgrad: operands must be 16-byte aligned");
  if (n == 0) {
    GODE_CHECK_CUDA(cudaMemset2DAsync(out, sizeof(float) * ldo, 0, sizeof(float) * ncols, d, st));
    GODE_CHECK_CUDA(cudaMemsetAsync(cs, 0, sizeof(float) * ncols, st));
    return GODE_OK;
  }
  const int64_t n_chunks = (n + 31) / 32;
  int grid = static_cast<int>(n_chunks < persistent_ctas() ? n_chunks : persistent_ctas());
  if (grid < 1) grid = 1;
  if (ws_bytes < sizeof(float) * (size_t)grid * (D + 1) * D) {
    set_error("gn_wgrad: workspace too small");
    return GODE_EWORKSPACE;
  }
  float* cs_partial = ws + (size_t)grid * D * D;
  constexpr size_t smem = 2 * 4 * (size_t)(D / 8) * 1168 + 128 + (size_t)D * D * 4;
  GODE_CHECK_CUDA(cudaFuncSetAttribute(k_wgrad_tc<D, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int passes = precision == GODE_PREC_TF32 ? 1 : 3;
  for (int c0 = 0; c0 < ncols; c0 += D) {
    const int gcols = ncols - c0 < D ? ncols - c0 : D;
    k_wgrad_tc<D, 4><<<grid, tc::WG_THREADS, smem, st>>>(n, y, G + c0, ws, cs_partial, eps, passes, 8, 0, ldg, gcols);
    GODE_LAUNCH_CHECK();
    k_wgrad_finish_raw<<<((D + 1) * gcols + 255) / 256, 256, 0, st>>>(grid, D, gcols, ws, cs_partial, out + c0, ldo, cs + c0);
    GODE_LAUNCH_CHECK();
  }
  return GODE_OK;
}

// GODE_TC_ACC: accuracy configuration of the 3xTF32 products of k_rows_ws (bit field, see the kernel).  Read at every
// launch so that one process can compare configurations (tools/transform_accuracy.py).
static int rows_acc_cfg(int mode) {
  const char* e = getenv("GODE_TC_ACC");
  if (e) return atoi(e);
  // MODE 0 (the transform) decides ReLU masks: the library-grade configuration.  MODE 1 (input gradient) does not.
  return mode == 0 ? GODE_TC_ACC_DEFAULT : 3;
}

// GODE_ROWS_WS: 1 (default) = warp-specialised k_rows_ws for d = 128, 0 = the single-pipeline k_rows_tc
static bool rows_ws_enabled() {
  static const int m = [] {
    const char* e = getenv("GODE_ROWS_WS");
    return e ? atoi(e) : 1;
  }();
  return m != 0;
}

template <int CPG, int MODE>
static int launch_rows_ws(int64_t n_rows, const float* X, float* Out, const float* W, const float* gamma, const float* beta,
                          float t, float eps, int passes, cudaStream_t st, const gode_push_route_t& pr, int grid,
                          int64_t ldx = 128, int64_t ldo = 128, int64_t ldw = 128) {
  constexpr int D = 128;
  constexpr size_t smem = 2 * (size_t)D * D * 4 + 4 * (size_t)128 * 32 * 4 + (size_t)128 * (D / 2) * 4 + D * 4 + 128;
  static bool configured = false;
  if (!configured) {
    GODE_CHECK_CUDA(cudaFuncSetAttribute(k_rows_ws<D, CPG, MODE, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GODE_CHECK_CUDA(cudaFuncSetAttribute(k_rows_ws<D, CPG, MODE, 19>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GODE_CHECK_CUDA(cudaFuncSetAttribute(k_rows_ws<D, CPG, MODE, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  static const int pf = [] {
    const char* e = getenv("GODE_TC_PREFETCH");   // L2 prefetch distance in tiles per CTA (default 0 = off: measured
    return e ? atoi(e) : 0;                       // 2.28 ms without vs 2.39 ms with, N = 10 M; the loads are not the limit)
  }();
  const int cfg = rows_acc_cfg(MODE);
  if (cfg == 3) k_rows_ws<D, CPG, MODE, 3><<<grid, tc::WS_THREADS, smem, st>>>(n_rows, X, Out, W, gamma, beta, t, eps, passes, pr, pf, cfg, ldx, ldo, ldw);
  else if (cfg == 19) k_rows_ws<D, CPG, MODE, 19><<<grid, tc::WS_THREADS, smem, st>>>(n_rows, X, Out, W, gamma, beta, t, eps, passes, pr, pf, cfg, ldx, ldo, ldw);
  else k_rows_ws<D, CPG, MODE, -1><<<grid, tc::WS_THREADS, smem, st>>>(n_rows, X, Out, W, gamma, beta, t, eps, passes, pr, pf, cfg, ldx, ldo, ldw);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

template <int D, int CPG, int MODE>
static int launch_rows_tc(int64_t n_rows, const float* X, float* Out, const float* W, const float* gamma, const float* beta,
                          float t, float eps, int passes, cudaStream_t st, const gode_push_route_t* push = nullptr) {
  constexpr size_t smem = 2 * (size_t)D * D * 4 + 2 * (size_t)128 * (D / 2) * 4 + D * 4 + 64;
  static bool configured = false;
  if (!configured) {
    GODE_CHECK_CUDA(cudaFuncSetAttribute(k_rows_tc<D, CPG, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int64_t n_tiles = (n_rows + 127) / 128;
  if (n_tiles == 0) return GODE_OK;
  const int grid = static_cast<int>(n_tiles < persistent_ctas() ? n_tiles : persistent_ctas());
  gode_push_route_t pr;
  if (push) pr = *push;
  else memset(&pr, 0, sizeof(pr));
  if (D == 128 && rows_ws_enabled()) return launch_rows_ws<CPG, MODE>(n_rows, X, Out, W, gamma, beta, t, eps, passes, st, pr, grid);
  k_rows_tc<D, CPG, MODE><<<grid, tc::THREADS, smem, st>>>(n_rows, X, Out, W, gamma, beta, t, eps, passes, pr);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

static bool rows_tc_shape(const gode_gcn_odefunc_t* f) {
  const int cpg = f->d / f->groups;
  return (f->d == 128 && cpg == 4) || (f->d == 64 && cpg == 2);
}
bool transform_tc_supported(const gode_gcn_odefunc_t* f) { return (tc_mask() & 1) && rows_tc_shape(f); }
bool input_grad_tc_supported(const gode_gcn_odefunc_t* f) { return (tc_mask() & 2) && rows_tc_shape(f); }

int transform_tc(const gode_gcn_odefunc_t* f, const float* y, float t, float* S, cudaStream_t st, int64_t row0, int64_t n_rows) {
  GODE_REQUIRE(al16(y) && al16(S), "transform_tc: operands must be 16-byte aligned");
  const int passes = f->precision == GODE_PREC_TF32 ? 1 : 3;
  const gode_push_route_t* push = f->push_S.ptr ? &f->push_S : nullptr;
  if (n_rows < 0) {        // whole block
    row0 = 0;
    n_rows = f->A.n_rows;
  }
  // rows beyond the owned block are the halo rows of a partitioned block (y and S then have A.n_cols rows)
  GODE_REQUIRE(row0 >= 0 && row0 + n_rows <= (f->A.n_cols > f->A.n_rows ? f->A.n_cols : f->A.n_rows),
               "transform_tc: row range outside the block");
  GODE_REQUIRE(!push || row0 == 0, "transform_tc: a fused push addresses rows from 0");
  y += row0 * f->d;
  S += row0 * f->d;
  if (f->d == 128) return launch_rows_tc<128, 4, 0>(n_rows, y, S, f->W, f->gamma, f->beta, t, f->gn_eps, passes, st, push);
  if (f->d == 64) return launch_rows_tc<64, 2, 0>(n_rows, y, S, f->W, f->gamma, f->beta, t, f->gn_eps, passes, st, push);
  set_error("transform_tc: unsupported width %d", f->d);
  return GODE_EINVAL;
}

// dz = dS * W[1:,:]^T
int input_grad_tc(const gode_gcn_odefunc_t* f, const float* gS, float* gz, cudaStream_t st) {
  GODE_REQUIRE(al16(gS) && al16(gz), "input_grad_tc: operands must be 16-byte aligned");
  const int passes = f->precision == GODE_PREC_TF32 ? 1 : 3;
  if (f->d == 128) return launch_rows_tc<128, 0, 1>(f->A.n_rows, gS, gz, f->W, f->gamma, f->beta, 0.f, f->gn_eps, passes, st);
  if (f->d == 64) return launch_rows_tc<64, 0, 1>(f->A.n_rows, gS, gz, f->W, f->gamma, f->beta, 0.f, f->gn_eps, passes, st);
  set_error("input_grad_tc: unsupported width %d", f->d);
  return GODE_EINVAL;
}


// ------------------------------------------------------------------------------------------------
// Out[r, c] = sum_k GroupNorm(y)[r, k] * W[1 + k, c] + t * W[0, c]  for a layer wider than d (the GAT node projection,
// 2 * heads * oh + 2 * heads columns): 128-column blocks on k_rows_ws (strided W / Out), the remaining 4 / 8 / 16 / 32
// attention-logit columns on a warp-per-row SIMT kernel (0.3 % of the flops; K = 128 in fp32 FMAs).
// ------------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(256) k_gn_cols(int64_t n_rows, const float* __restrict__ Y, const float* __restrict__ W /*row 0 = t row*/,
                                                 int64_t ldw, const float* __restrict__ gamma, const float* __restrict__ beta, float t,
                                                 float eps, float* __restrict__ Out, int64_t ldo) {
  constexpr int D = 128;
  // gamma[k] * W[1 + k, c] at (k / 4) * LS + (k % 4) * NC + c: a lane reads the four rows of its group as float4s, and
  // LS / 4 odd spreads the lanes of a quarter warp over all eight 16-byte bank groups (LS = 4 * NC put every lane on the
  // same banks: 32-way conflicts, 2.15 ms per launch at N = 1 M)
  constexpr int LS = 4 * NC + 4;
  __shared__ __align__(16) float sW[(D / 4) * LS];
  __shared__ float sR[NC];                       // t * W[0, c] + sum_k beta[k] W[1 + k, c]
  for (int i = threadIdx.x; i < D * NC; i += blockDim.x) {
    const int k = i / NC, c = i % NC;
    sW[(k >> 2) * LS + (k & 3) * NC + c] = __ldg(gamma + k) * __ldg(W + (int64_t)(1 + k) * ldw + c);
  }
  if (threadIdx.x < NC) {
    const int c = threadIdx.x;
    float r = t * __ldg(W + c);
    for (int k = 0; k < D; ++k) r = fmaf(__ldg(beta + k), __ldg(W + (int64_t)(1 + k) * ldw + c), r);
    sR[c] = r;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // the row loop is latency-bound (one 512-byte load, ~100 dependent instructions, one store per row): the next two rows'
  // loads are issued before this row's arithmetic
  float4 nx0 = make_float4(0.f, 0.f, 0.f, 0.f), nx1 = nx0;
  if (warp0 < n_rows) nx0 = __ldcs(reinterpret_cast<const float4*>(Y + warp0 * D + lane * 4));
  if (warp0 + nwarps < n_rows) nx1 = __ldcs(reinterpret_cast<const float4*>(Y + (warp0 + nwarps) * D + lane * 4));
  for (int64_t r = warp0; r < n_rows; r += nwarps) {
    const float4 x = tc::normalize4<4>(nx0, eps);   // lane = one group
    nx0 = nx1;
    if (r + 2 * nwarps < n_rows) nx1 = __ldcs(reinterpret_cast<const float4*>(Y + (r + 2 * nwarps) * D + lane * 4));
    float acc[NC];
#pragma unroll
    for (int c = 0; c < NC; c += 4) {
      const float4 w0 = *reinterpret_cast<const float4*>(sW + lane * LS + 0 * NC + c);
      const float4 w1 = *reinterpret_cast<const float4*>(sW + lane * LS + 1 * NC + c);
      const float4 w2 = *reinterpret_cast<const float4*>(sW + lane * LS + 2 * NC + c);
      const float4 w3 = *reinterpret_cast<const float4*>(sW + lane * LS + 3 * NC + c);
      acc[c + 0] = x.x * w0.x + x.y * w1.x + x.z * w2.x + x.w * w3.x;
      acc[c + 1] = x.x * w0.y + x.y * w1.y + x.z * w2.y + x.w * w3.y;
      acc[c + 2] = x.x * w0.z + x.y * w1.z + x.z * w2.z + x.w * w3.z;
      acc[c + 3] = x.x * w0.w + x.y * w1.w + x.z * w2.w + x.w * w3.w;
    }
    // transposed butterfly: every exchange halves the number of columns a lane still carries (NC - 1 + log2(32 / NC)
    // shuffles instead of 5 * NC); lane bits, from the top, select the column a lane ends up with
    int col = 0;
    bool writer = true;
    {
      int nv = NC;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        if (nv > 1) {
          const int half = nv / 2;
          const bool up = (lane & o) != 0;
#pragma unroll
          for (int i = 0; i < NC / 2; ++i)
            if (i < half) {
              const float send = up ? acc[i] : acc[i + half];
              const float recv = __shfl_xor_sync(0xffffffffu, send, o);
              acc[i] = (up ? acc[i + half] : acc[i]) + recv;
            }
          if (up) col += half;
          nv = half;
        } else {
          acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], o);
          writer = writer && (lane & o) == 0;
        }
      }
    }
    if (writer) Out[r * ldo + col] = acc[0] + sR[col];
  }
}

bool gn_linear_tc_supported(int d, int groups, int ncols) {
  const int rem = ncols % 128;
  return (tc_mask() & 1) && rows_ws_enabled() && d == 128 && groups == 32 && ncols >= 4 &&
         (rem == 0 || rem == 4 || rem == 8 || rem == 16 || rem == 32);
}

int gn_linear_tc(int64_t n, int d, int groups, float eps, const float* y, const float* gamma, const float* beta, const float* W,
                 int64_t ldw, float t, int ncols, float* out, int64_t ldo, int precision, cudaStream_t st) {
  GODE_REQUIRE(gn_linear_tc_supported(d, groups, ncols), "gn_linear: d = 128 with 32 groups; columns = k * 128 + {0, 4, 8, 16, 32}");
  GODE_REQUIRE(ldw >= ncols && ldo >= ncols && ldo % 4 == 0 && al16(y) && al16(out), "gn_linear: bad leading dimension / alignment");
  if (n == 0) return GODE_OK;
  const int passes = precision == GODE_PREC_TF32 ? 1 : 3;
  const int64_t n_tiles = (n + 127) / 128;
  const int grid = static_cast<int>(n_tiles < persistent_ctas() ? n_tiles : persistent_ctas());
  gode_push_route_t pr;
  memset(&pr, 0, sizeof(pr));
  int c0 = 0;
  for (; c0 + 128 <= ncols; c0 += 128) {
    int rc = launch_rows_ws<4, 0>(n, y, out + c0, W + c0, gamma, beta, t, eps, passes, st, pr, grid, 128, ldo, ldw);
    if (rc) return rc;
  }
  const int rem = ncols - c0;
  if (rem > 0) {
    const int64_t want = (n * 32 + 255) / 256;
    const int64_t cap = 8LL * sm_count();
    const unsigned g = static_cast<unsigned>(want < cap ? want : cap);
    float* o = out + c0;
    const float* w = W + c0;
    if (rem == 4) k_gn_cols<4><<<g, 256, 0, st>>>(n, y, w, ldw, gamma, beta, t, eps, o, ldo);
    else if (rem == 8) k_gn_cols<8><<<g, 256, 0, st>>>(n, y, w, ldw, gamma, beta, t, eps, o, ldo);
    else if (rem == 16) k_gn_cols<16><<<g, 256, 0, st>>>(n, y, w, ldw, gamma, beta, t, eps, o, ldo);
    else k_gn_cols<32><<<g, 256, 0, st>>>(n, y, w, ldw, gamma, beta, t, eps, o, ldo);
    GODE_LAUNCH_CHECK();
  }
  return GODE_OK;
}

// ------------------------------------------------------------------------------------------------
// k_gemm_tc: general  C[M, N] = act(A[M, K] * Bt[N, K]^T + bias)  on tcgen05, 3xTF32 -- the dense products that are
// not the d x d ones above: the QC edge encoder (QC/layers.py:76-86: [E, 2667] x [2667, 5329], 28.4 MFLOP per edge --
// SURVEY 8a "dominates QC"), the GAT node projections, the input layer of the GCN models.
//
// Round 2: first run on a B200 green (tests/test_gpu_gemm_tc.py, 14 shapes against fp64); ops.linear routes large products here.
//
// Both operands are K-major (K contiguous): a weight stored [K, N] is transposed once by the caller.  One persistent CTA
// per SM walks 128 x 128 output tiles (tile index = tm + tiles_m * tn, so concurrently running CTAs share one B tile);
// K is consumed in stages of 32 (one 128-byte swizzle atom): eight producer warps load the A and the B part of a stage
// (a quarter warp reads one 128-byte line), split hi / lo, store both into 128-byte-swizzled K-major stages of a 2-deep
// ring and arrive on the stage's "full" barrier; one thread issues 4 K-steps x 3 passes of tcgen05.mma per stage into one
// of two TMEM accumulator sets and commits the stage's "empty" barrier; four epilogue warps drain finished sets into
// shared-memory running sums and finish a tile (bias, ReLU, row stores).  Rows / columns / K beyond the matrix are zero-filled on load and masked on
// store, so M, N, K are arbitrary; 128-bit accesses need lda, ldb (ldc) % 4 == 0 and 16-byte aligned bases, otherwise the
// kernel falls back to 32-bit accesses.
// ------------------------------------------------------------------------------------------------
namespace tc {
constexpr int GT_STAGES = 2;   // operand stages (64 KB each: A_hi | A_lo | B_hi | B_lo of 128 x 32)
constexpr int GT_KG = 8;       // stages (K = 256 = 32 MMA K steps) accumulated inside the tensor core before a drain
}

// Accumulation is hierarchical, as in k_wgrad_tc: the MMA's own fp32 accumulate drops low bits at EVERY K step (measured:
// ~0.5 ulp per step, one-sided), which over the 334 K steps of the QC edge encoder (K = 2667) would be 2e-5 -- above the
// 1e-5 bar.  So GT_KG stages go into one of two TMEM accumulator sets (hi*hi and the correction terms apart), and the four
// epilogue warps add every finished set into fp32 running sums in shared memory (round-to-nearest adds) while the MMAs
// continue into the other set; after a tile's last group the sums get bias / ReLU and are stored as full rows.
// Output tile of work item t.  Tiles are visited in bands of `gm` row tiles, column tile by column tile inside a band, so
// that the ~148 tiles in flight at any time share ~gm A panels and ~148 / gm B panels (both L2 resident) instead of 148
// distinct A panels and one B panel: with the row-tile-fastest order every A panel of the QC edge encoder ([146 618, 2667],
// 1.56 GB) came back from HBM once per column tile (42 x 1.56 GB per product).
__device__ __forceinline__ void gemm_tile_of(int64_t t, int64_t tiles_m, int64_t tiles_n, int gm, int64_t& tm, int64_t& tn) {
  const int64_t per_band = (int64_t)gm * tiles_n;
  const int64_t band = t / per_band, within = t - band * per_band;
  const int64_t m0 = band * gm;
  const int64_t rows = tiles_m - m0 < gm ? tiles_m - m0 : gm;
  tn = within / rows;
  tm = m0 + (within - tn * rows);
}

__global__ void __launch_bounds__(tc::WS_THREADS, 1)
k_gemm_tc(int64_t M, int64_t N, int64_t K, const float* __restrict__ A, int64_t lda, const float* __restrict__ Bt, int64_t ldb,
          float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int relu, int passes, int vec_in, int vec_out,
          int k_splits /*> 1: the K range is cut into k_splits slabs, slab s writes its partial product to C + s * M * ldc*/,
          int band /*row tiles per band of the tile order, see gemm_tile_of*/, int pf_stages /*L2 prefetch distance of A, in stages*/) {
  using namespace tc;
  constexpr int BK = 32;                          // K per stage: one 128-byte swizzle atom of tf32
  constexpr int NI = 4;                           // warp-instructions per producer warp per operand per stage
  constexpr uint32_t T_BYTES = 128 * BK * 4;      // one of hi / lo of one operand of one stage (16 KB)
  constexpr uint32_t STAGE_BYTES = 4 * T_BYTES;   // A_hi | A_lo | B_hi | B_lo
  constexpr uint32_t SUM_BYTES = 128 * 128 * 4;   // running sums of one output tile
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sums = reinterpret_cast<float*>(smem + GT_STAGES * STAGE_BYTES);   // [128 rows][32 chunks of 4, XOR-swizzled by row]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(sums) + SUM_BYTES);
  // bars: [0, S) full, [S, 2S) empty, [2S, 2S+2) set_full, [2S+2, 2S+4) set_drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GT_STAGES + 4);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tiles_m = (M + 127) / 128, tiles_n = (N + 127) / 128;
  const int64_t mn_tiles = tiles_m * tiles_n;
  const int64_t n_tiles = mn_tiles * k_splits;    // work items: (output tile, K slab); every slab has the same number of stages
  if ((int64_t)blockIdx.x >= n_tiles) return;
  const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int64_t KS = ((K + BK - 1) / BK + k_splits - 1) / k_splits;   // stages per work item (the last slab is zero-padded)
  const int64_t n_steps = my_tiles * KS;

  if (tid == 0) {
    for (int s = 0; s < GT_STAGES; ++s) {
      mbar_init(&bars[s], WS_PRODUCERS);
      mbar_init(&bars[GT_STAGES + s], 1);
    }
    mbar_init(&bars[2 * GT_STAGES + 0], 1);
    mbar_init(&bars[2 * GT_STAGES + 1], 1);
    mbar_init(&bars[2 * GT_STAGES + 2], 128);
    mbar_init(&bars[2 * GT_STAGES + 3], 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);   // two accumulator sets x (hi*hi | correction terms)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t IDESC = make_idesc(128, 128, false, false);

  if (warp < 8) {
    // =============================== producers ============================================================
    // registers of one step: 4 chunks of A and 4 of B (row = (warp * 4 + i) * 4 + lane / 8, 16-byte chunk = lane % 8)
    float4 raw[2][2 * NI];
    // Load cursor.  The stages are loaded strictly in order, so the position (tile of this CTA, stage of the tile) is kept
    // in counters and the tile's coordinates / row pointers are recomputed once per tile: with the divisions
    // (step / KS, v % mn_tiles, the band order) evaluated per stage the producers executed 740 instructions per warp per
    // stage -- ncu round 2: 43 % issue-slot utilisation against 22 % tensor-pipe utilisation, 3 300 cycles per stage.
    int64_t ld_it = 0;                              // tile iteration of this CTA
    int64_t ld_ks = 0;                              // stage inside the tile
    bool ld_done = false;
    const float* pa[NI];
    const float* pb[NI];
    bool va[NI], vb[NI];
    int64_t kbase = 0;
    auto ld_setup = [&]() {
      const int64_t v = blockIdx.x + ld_it * (int64_t)gridDim.x;
      ld_done = v >= n_tiles;
      if (ld_done) return;
      const int64_t t = v % mn_tiles, sp = v / mn_tiles;
      int64_t tm, tn;
      gemm_tile_of(t, tiles_m, tiles_n, band, tm, tn);
      kbase = sp * KS * BK + (lane & 7) * 4;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int rr = (warp * NI + i) * 4 + (lane >> 3);
        const int64_t ra = tm * 128 + rr, rb = tn * 128 + rr;
        va[i] = ra < M;
        vb[i] = rb < N;
        pa[i] = A + (va[i] ? ra : 0) * lda;
        pb[i] = Bt + (vb[i] ? rb : 0) * ldb;
      }
    };
    ld_setup();
    // one 16-byte chunk of A (slot i) and of B (slot NI + i) of the stage at the cursor
    auto load_slot = [&](float4 (&r)[2 * NI], int i) {
      if (ld_done) return;
      const int64_t k = kbase + ld_ks * BK;
      const bool kin = k < K, ktail = k + 3 >= K;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float* __restrict__ src = (h == 0 ? pa[i] : pb[i]) + k;
        const bool ok = kin && (h == 0 ? va[i] : vb[i]);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) {
          if (vec_in) {
            v = __ldg(reinterpret_cast<const float4*>(src));     // in bounds of the padded row (ld % 4 == 0, ld >= K)
            if (ktail) {
              if (k + 1 >= K) v.y = 0.f;
              if (k + 2 >= K) v.z = 0.f;
              v.w = 0.f;
            }
          } else {
            v.x = __ldg(src);
            if (k + 1 < K) v.y = __ldg(src + 1);
            if (k + 2 < K) v.z = __ldg(src + 2);
            if (k + 3 < K) v.w = __ldg(src + 3);
          }
          // ask L2 for this row's line `pf_stages` stages ahead (no registers; one request per 128-byte line)
          if (h == 0 && pf_stages > 0 && (lane & 7) == 0 && k + pf_stages * BK < K) prefetch_l2(src + pf_stages * BK);
        }
        r[h * NI + i] = v;
      }
    };
    auto ld_advance = [&]() {
      if (ld_done) return;
      if (++ld_ks == KS) {
        ld_ks = 0;
        ++ld_it;
        ld_setup();
      }
    };
    auto load_raw = [&](float4 (&r)[2 * NI]) {
#pragma unroll
      for (int i = 0; i < NI; ++i) load_slot(r, i);
      ld_advance();
    };
    auto do_step = [&](float4 (&r)[2 * NI], int64_t step) {
      if (step >= n_steps) return;
      const int s = static_cast<int>(step % GT_STAGES);
      const int64_t use = step / GT_STAGES;
      if (use >= 1) mbar_wait(&bars[GT_STAGES + s], static_cast<uint32_t>((use - 1) & 1));
      unsigned char* base = smem + (size_t)s * STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int rr = (warp * NI + i) * 4 + (lane >> 3);
        const uint32_t off = (rr >> 3) * 1024 + (rr & 7) * 128 + (((lane & 7) ^ (rr & 7)) << 4);
        float4 hi, lo;
        split4(r[i], hi, lo);
        *reinterpret_cast<float4*>(base + off) = hi;
        *reinterpret_cast<float4*>(base + T_BYTES + off) = lo;
        split4(r[NI + i], hi, lo);
        *reinterpret_cast<float4*>(base + 2 * T_BYTES + off) = hi;
        *reinterpret_cast<float4*>(base + 3 * T_BYTES + off) = lo;
      }
      load_raw(r);     // the stage two ahead, into the registers just consumed (issuing these chunk by chunk between the
                       // stores above measured SLOWER: 47 vs 36 ms on the QC forward product)
      fence_async_smem();
      mbar_arrive(&bars[s]);
    };
    load_raw(raw[0]);
    load_raw(raw[1]);
    for (int64_t step = 0; step < n_steps; step += 2) {
      do_step(raw[0], step);
      do_step(raw[1], step + 1);
    }
  } else if (warp == 12) {
    // =============================== MMA issue (one thread) ===============================================
    if (lane == 0) {
      int64_t g = 0;                                 // accumulation groups issued so far (over all tiles of this CTA)
      int64_t ks = -1;                               // stage inside the tile (a counter: no division per stage)
      for (int64_t step = 0; step < n_steps; ++step) {
        const int s = static_cast<int>(step % GT_STAGES);
        if (++ks == KS) ks = 0;
        const int set = static_cast<int>(g & 1);
        const bool first = ks % GT_KG == 0;          // first stage of its group
        if (first && g >= 2) mbar_wait(&bars[2 * GT_STAGES + 2 + set], static_cast<uint32_t>(((g >> 1) - 1) & 1));
        mbar_wait(&bars[s], static_cast<uint32_t>((step / GT_STAGES) & 1));
        tc_fence_after();
        const uint32_t base = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint32_t tmem_d = tmem_base + set * 256;
#pragma unroll
        for (int q = 0; q < BK / 8; ++q) {
          const uint64_t a_hi = make_desc_sw128(base + q * 32), a_lo = make_desc_sw128(base + T_BYTES + q * 32);
          const uint64_t b_hi = make_desc_sw128(base + 2 * T_BYTES + q * 32), b_lo = make_desc_sw128(base + 3 * T_BYTES + q * 32);
          const uint32_t acc = (first && q == 0) ? 0u : 1u;
          mma_tf32(tmem_d, a_hi, b_hi, IDESC, acc);
          if (passes == 3) {
            mma_tf32(tmem_d + 128, a_lo, b_hi, IDESC, acc);
            mma_tf32(tmem_d + 128, a_hi, b_lo, IDESC, 1u);
          }
        }
        mma_commit(&bars[GT_STAGES + s]);
        if (ks % GT_KG == GT_KG - 1 || ks == KS - 1) {
          mma_commit(&bars[2 * GT_STAGES + set]);    // this group's partial product is complete
          ++g;
        }
      }
    }
  } else {
    // =============================== drain + epilogue (warps 8-11) ========================================
    constexpr int CH = 32;                           // 16-byte chunks per row of the sums tile
    const int q = warp & 3;
    const int et = tid - WS_PRODUCERS;
    const int row = q * 32 + lane;                   // TMEM lane = tile row; a thread adds into its own row only
    const int64_t n_groups = (KS + GT_KG - 1) / GT_KG;
    int64_t g = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t v = blockIdx.x + it * (int64_t)gridDim.x;
      const int64_t t = v % mn_tiles, sp = v / mn_tiles;
      int64_t tm, tn;
      gemm_tile_of(t, tiles_m, tiles_n, band, tm, tn);
      float* __restrict__ Cs = C + sp * M * ldc;
      for (int64_t gi = 0; gi < n_groups; ++gi, ++g) {
        const int set = static_cast<int>(g & 1);
        mbar_wait(&bars[2 * GT_STAGES + set], static_cast<uint32_t>((g >> 1) & 1));
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + set * 256 + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
        for (int cb = 0; cb < 128; cb += 32) {
          float v[32];
          tmem_ld32(tmem_d + cb, v);
          if (passes == 3) {
            float w[32];
            tmem_ld32(tmem_d + 128 + cb, w);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += w[j];
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4* p = reinterpret_cast<float4*>(sums + row * 128 + ((((cb + j) >> 2) ^ (row & (CH - 1))) << 2));
            float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (gi > 0) {
              const float4 old = *p;
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *p = o;
          }
        }
        tc_fence_before();
        mbar_arrive(&bars[2 * GT_STAGES + 2 + set]);
      }
      bar_sync_named(2, 128);                        // every row of the tile's sums is final
#pragma unroll 4
      for (int i = 0; i < 32; ++i) {
        const int idx = i * 128 + et;
        const int r = idx / CH, c = idx % CH;
        const int64_t grow = tm * 128 + r;
        const int64_t col = tn * 128 + c * 4;
        if (grow < M && col < N) {
          float4 o = *reinterpret_cast<const float4*>(sums + r * 128 + ((c ^ (r & (CH - 1))) << 2));
          if (bias) {
            o.x += __ldg(bias + col);
            if (col + 1 < N) o.y += __ldg(bias + col + 1);
            if (col + 2 < N) o.z += __ldg(bias + col + 2);
            if (col + 3 < N) o.w += __ldg(bias + col + 3);
          }
          if (relu) {
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
          }
          float* dst = Cs + grow * ldc + col;
          if (vec_out && col + 3 < N) {
            *reinterpret_cast<float4*>(dst) = o;
          } else {
            dst[0] = o.x;
            if (col + 1 < N) dst[1] = o.y;
            if (col + 2 < N) dst[2] = o.z;
            if (col + 3 < N) dst[3] = o.w;
          }
        }
      }
      bar_sync_named(2, 128);                        // the next tile's first group overwrites the sums
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// k_gemm_tc2: the same product for 16-byte-aligned operands (lda, ldb % 4 == 0), with the operand loads taken out of the
// registers.  k_gemm_tc keeps two stages of raw operands in registers: a load is issued one stage period before it is used,
// so the period cannot drop below the loaded L2 latency (ncu round 2: 2 350 cycles per stage against ~750 of tensor work).
// Here every producer thread copies its own 16-byte chunks of the next GT2_RING stages into a raw ring in shared memory
// with cp.async (thread-private slots: no barrier, cp.async.wait_group orders a thread's own copies; out-of-range rows and
// the K tail are zero-filled by the copy's src-size operand), reads them back, splits and stores the hi / lo MMA stages.
// Room for the ring comes from moving the running sums out of shared memory: they live in a fourth TMEM accumulator that
// the epilogue warps update with tcgen05.ld -> fp32 add -> tcgen05.st (TMEM: hi*hi set 0 | hi*hi set 1 | correction terms
// | running sums, 128 columns each; the correction terms are 2^-11 of the product and accumulate over the whole tile).
// ------------------------------------------------------------------------------------------------
namespace tc {
constexpr int GT2_RING = 3;    // raw stages in flight per CTA (32 KB each)

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
}  // namespace tc

__global__ void __launch_bounds__(tc::WS_THREADS, 1)
k_gemm_tc2(int64_t M, int64_t N, int64_t K, const float* __restrict__ A, int64_t lda, const float* __restrict__ Bt, int64_t ldb,
           float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int relu, int passes, int vec_out, int k_splits,
           int band) {
  using namespace tc;
  constexpr int BK = 32;
  constexpr int NI = 4;
  constexpr uint32_t T_BYTES = 128 * BK * 4;      // one of hi / lo of one operand of one stage (16 KB)
  constexpr uint32_t STAGE_BYTES = 4 * T_BYTES;   // A_hi | A_lo | B_hi | B_lo
  constexpr uint32_t RAW_BYTES = 2 * T_BYTES;     // raw A | raw B of one stage (32 KB)
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* ring = smem + GT_STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + GT2_RING * RAW_BYTES);
  // bars: [0, S) full, [S, 2S) empty, [2S, 2S+2) set_full, [2S+2, 2S+4) set_drained, [2S+4] tile_drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GT_STAGES + 5);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tiles_m = (M + 127) / 128, tiles_n = (N + 127) / 128;
  const int64_t mn_tiles = tiles_m * tiles_n;
  const int64_t n_tiles = mn_tiles * k_splits;
  if ((int64_t)blockIdx.x >= n_tiles) return;
  const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int64_t KS = ((K + BK - 1) / BK + k_splits - 1) / k_splits;
  const int64_t n_steps = my_tiles * KS;
  const int64_t n_groups = (KS + GT_KG - 1) / GT_KG;

  if (tid == 0) {
    for (int s = 0; s < GT_STAGES; ++s) {
      mbar_init(&bars[s], WS_PRODUCERS);
      mbar_init(&bars[GT_STAGES + s], 1);
    }
    mbar_init(&bars[2 * GT_STAGES + 0], 1);
    mbar_init(&bars[2 * GT_STAGES + 1], 1);
    mbar_init(&bars[2 * GT_STAGES + 2], 128);
    mbar_init(&bars[2 * GT_STAGES + 3], 128);
    mbar_init(&bars[2 * GT_STAGES + 4], 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t IDESC = make_idesc(128, 128, false, false);
  constexpr uint32_t T_CORR = 256, T_SUM = 384;

  if (warp < 8) {
    // =============================== producers ============================================================
    int64_t ld_it = 0, ld_ks = 0;                   // load cursor: tile iteration of this CTA, stage inside the tile
    bool ld_done = false;
    const float* pa[NI];
    const float* pb[NI];
    bool va[NI], vb[NI];
    int64_t kbase = 0;
    auto ld_setup = [&]() {
      const int64_t v = blockIdx.x + ld_it * (int64_t)gridDim.x;
      ld_done = v >= n_tiles;
      if (ld_done) return;
      const int64_t t = v % mn_tiles, sp = v / mn_tiles;
      int64_t tm, tn;
      gemm_tile_of(t, tiles_m, tiles_n, band, tm, tn);
      kbase = sp * KS * BK + (lane & 7) * 4;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int rr = (warp * NI + i) * 4 + (lane >> 3);
        const int64_t ra = tm * 128 + rr, rb = tn * 128 + rr;
        va[i] = ra < M;
        vb[i] = rb < N;
        pa[i] = A + (va[i] ? ra : 0) * lda;
        pb[i] = Bt + (vb[i] ? rb : 0) * ldb;
      }
    };
    ld_setup();
    const uint32_t ring_u32 = smem_u32(ring) + tid * 16;    // chunk i of slot r: ring + r * RAW_BYTES + i * 4096 + tid * 16
    // copies of the stage at the cursor into ring slot `slot` (nothing when the CTA's work is exhausted); always one group
    auto issue = [&](int slot) {
      if (!ld_done) {
        const int64_t k = kbase + ld_ks * BK;
        const int64_t left = K - k;                 // valid floats from k on (<= 0: whole chunk zero-filled)
        const uint32_t kb = left >= 4 ? 16u : (left > 0 ? static_cast<uint32_t>(left) * 4u : 0u);
        const int64_t ko = left > 0 ? k : 0;        // keep the address inside the row when nothing is read
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          cp_async16(ring_u32 + slot * RAW_BYTES + i * 4096, pa[i] + ko, va[i] ? kb : 0u);
          cp_async16(ring_u32 + slot * RAW_BYTES + (NI + i) * 4096, pb[i] + ko, vb[i] ? kb : 0u);
        }
        if (++ld_ks == KS) {
          ld_ks = 0;
          ++ld_it;
          ld_setup();
        }
      }
      cp_async_commit();
    };
#pragma unroll
    for (int r = 0; r < GT2_RING; ++r) issue(r);
    int slot = 0;
    for (int64_t step = 0; step < n_steps; ++step) {
      const int s = static_cast<int>(step % GT_STAGES);
      const int64_t use = step / GT_STAGES;
      cp_async_wait<GT2_RING - 1>();                // this thread's copies of stage `step` have landed
      float4 raw[2 * NI];
      const unsigned char* rs = ring + (size_t)slot * RAW_BYTES + tid * 16;
#pragma unroll
      for (int i = 0; i < 2 * NI; ++i) raw[i] = *reinterpret_cast<const float4*>(rs + i * 4096);
      if (use >= 1) mbar_wait(&bars[GT_STAGES + s], static_cast<uint32_t>((use - 1) & 1));
      unsigned char* base = smem + (size_t)s * STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int rr = (warp * NI + i) * 4 + (lane >> 3);
        const uint32_t off = (rr >> 3) * 1024 + (rr & 7) * 128 + (((lane & 7) ^ (rr & 7)) << 4);
        float4 hi, lo;
        split4(raw[i], hi, lo);
        *reinterpret_cast<float4*>(base + off) = hi;
        *reinterpret_cast<float4*>(base + T_BYTES + off) = lo;
        split4(raw[NI + i], hi, lo);
        *reinterpret_cast<float4*>(base + 2 * T_BYTES + off) = hi;
        *reinterpret_cast<float4*>(base + 3 * T_BYTES + off) = lo;
      }
      issue(slot);                                  // stage step + GT2_RING into the slot just read (values are in registers)
      fence_async_smem();
      mbar_arrive(&bars[s]);
      if (++slot == GT2_RING) slot = 0;
    }
    cp_async_wait<0>();
  } else if (warp == 12) {
    // =============================== MMA issue (one thread) ===============================================
    if (lane == 0) {
      int64_t g = 0;                                 // accumulation groups issued so far (over all tiles of this CTA)
      int64_t ks = -1;
      int64_t tile_it = 0;
      for (int64_t step = 0; step < n_steps; ++step) {
        const int s = static_cast<int>(step % GT_STAGES);
        if (++ks == KS) {
          ks = 0;
          ++tile_it;
        }
        const int set = static_cast<int>(g & 1);
        const bool first = ks % GT_KG == 0;          // first stage of its group
        if (first && g >= 2) mbar_wait(&bars[2 * GT_STAGES + 2 + set], static_cast<uint32_t>(((g >> 1) - 1) & 1));
        // the correction accumulator and the running sums belong to one tile at a time
        if (ks == 0 && tile_it >= 1) mbar_wait(&bars[2 * GT_STAGES + 4], static_cast<uint32_t>((tile_it - 1) & 1));
        mbar_wait(&bars[s], static_cast<uint32_t>((step / GT_STAGES) & 1));
        tc_fence_after();
        const uint32_t base = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint32_t tmem_d = tmem_base + set * 128;
#pragma unroll
        for (int q = 0; q < BK / 8; ++q) {
          const uint64_t a_hi = make_desc_sw128(base + q * 32), a_lo = make_desc_sw128(base + T_BYTES + q * 32);
          const uint64_t b_hi = make_desc_sw128(base + 2 * T_BYTES + q * 32), b_lo = make_desc_sw128(base + 3 * T_BYTES + q * 32);
          mma_tf32(tmem_d, a_hi, b_hi, IDESC, (first && q == 0) ? 0u : 1u);
          if (passes == 3) {
            mma_tf32(tmem_base + T_CORR, a_lo, b_hi, IDESC, (ks == 0 && q == 0) ? 0u : 1u);
            mma_tf32(tmem_base + T_CORR, a_hi, b_lo, IDESC, 1u);
          }
        }
        mma_commit(&bars[GT_STAGES + s]);
        if (ks % GT_KG == GT_KG - 1 || ks == KS - 1) {
          mma_commit(&bars[2 * GT_STAGES + set]);
          ++g;
        }
      }
    }
  } else {
    // =============================== drain + epilogue (warps 8-11) ========================================
    const int q = warp & 3;
    const int row = q * 32 + lane;                   // TMEM lane = tile row = the row this thread stores
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    int64_t g = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t v = blockIdx.x + it * (int64_t)gridDim.x;
      const int64_t t = v % mn_tiles, sp = v / mn_tiles;
      int64_t tm, tn;
      gemm_tile_of(t, tiles_m, tiles_n, band, tm, tn);
      float* __restrict__ Cs = C + sp * M * ldc;
      const int64_t grow = tm * 128 + row;
      for (int64_t gi = 0; gi < n_groups; ++gi, ++g) {
        const int set = static_cast<int>(g & 1);
        const bool last = gi == n_groups - 1;
        mbar_wait(&bars[2 * GT_STAGES + set], static_cast<uint32_t>((g >> 1) & 1));
        tc_fence_after();
        const uint32_t t_main = tmem_base + set * 128 + lane_off;
#pragma unroll 1
        for (int cb = 0; cb < 128; cb += 32) {
          float acc[32];
          tmem_ld32(t_main + cb, acc);
          if (gi > 0) {
            float w[32];
            tmem_ld32(tmem_base + T_SUM + lane_off + cb, w);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] += w[j];
          }
          if (!last) {
            tmem_st32(tmem_base + T_SUM + lane_off + cb, acc);
            continue;
          }
          if (passes == 3) {
            float w[32];
            tmem_ld32(tmem_base + T_CORR + lane_off + cb, w);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] += w[j];
          }
          if (cb == 96) {                            // every accumulator of the tile has been read out
            tc_fence_before();
            mbar_arrive(&bars[2 * GT_STAGES + 2 + set]);
            mbar_arrive(&bars[2 * GT_STAGES + 4]);
          }
          if (grow < M) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int64_t col = tn * 128 + cb + j;
              if (col >= N) break;
              float4 o = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
              if (bias) {
                o.x += __ldg(bias + col);
                if (col + 1 < N) o.y += __ldg(bias + col + 1);
                if (col + 2 < N) o.z += __ldg(bias + col + 2);
                if (col + 3 < N) o.w += __ldg(bias + col + 3);
              }
              if (relu) {
                o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
              }
              float* dst = Cs + grow * ldc + col;
              if (vec_out && col + 3 < N) {
                *reinterpret_cast<float4*>(dst) = o;
              } else {
                dst[0] = o.x;
                if (col + 1 < N) dst[1] = o.y;
                if (col + 2 < N) dst[2] = o.z;
                if (col + 3 < N) dst[3] = o.w;
              }
            }
          }
        }
        if (!last) {
          tc_fence_before();
          mbar_arrive(&bars[2 * GT_STAGES + 2 + set]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int gemm_tc(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* Bt, int64_t ldb, float* C, int64_t ldc,
            const float* bias, int relu, int precision, cudaStream_t st, int k_splits) {
  GODE_REQUIRE(M >= 0 && N >= 0 && K >= 1 && lda >= K && ldb >= K && ldc >= N, "gemm_tc: bad shape");
  GODE_REQUIRE(k_splits >= 1 && (k_splits == 1 || (!bias && !relu)), "gemm_tc: a split-K product has no bias / ReLU epilogue");
  if (M == 0 || N == 0) return GODE_OK;
  GODE_REQUIRE(A && Bt && C, "gemm_tc: null pointer");
  constexpr size_t smem = tc::GT_STAGES * 4 * (size_t)128 * 32 * 4 + (size_t)128 * 128 * 4 + 256;
  static bool configured = false;
  if (!configured) {
    GODE_CHECK_CUDA(cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int64_t n_tiles = ((M + 127) / 128) * ((N + 127) / 128) * k_splits;
  const int grid = static_cast<int>(n_tiles < persistent_ctas() ? n_tiles : persistent_ctas());
  const int vec_in = (lda % 4 == 0) && (ldb % 4 == 0) && al16(A) && al16(Bt);
  const int vec_out = (ldc % 4 == 0) && al16(C) && (k_splits == 1 || (M * ldc) % 4 == 0);
  // GODE_GEMM_BAND: row tiles per band of the tile order (default 16: 16 A panels + ~9 B panels in flight)
  const char* be = getenv("GODE_GEMM_BAND");
  int band = be ? atoi(be) : 16;
  if (band < 1) band = 1;
  const char* pe = getenv("GODE_GEMM_PREFETCH");    // L2 prefetch distance of the A operand in stages of 32 columns (0 = off)
  const int pf_stages = pe ? atoi(pe) : 0;
  // GODE_GEMM_V: 2 (default) = k_gemm_tc2 (cp.async raw ring, TMEM running sums) for aligned operands, 1 = k_gemm_tc
  const char* ve = getenv("GODE_GEMM_V");
  if (vec_in && (!ve || atoi(ve) != 1)) {
    constexpr size_t smem2 = tc::GT_STAGES * 4 * (size_t)128 * 32 * 4 + tc::GT2_RING * 2 * (size_t)128 * 32 * 4 + 256;
    static bool configured2 = false;
    if (!configured2) {
      GODE_CHECK_CUDA(cudaFuncSetAttribute(k_gemm_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      configured2 = true;
    }
    k_gemm_tc2<<<grid, tc::WS_THREADS, smem2, st>>>(M, N, K, A, lda, Bt, ldb, C, ldc, bias, relu,
                                                    precision == GODE_PREC_TF32 ? 1 : 3, vec_out, k_splits, band);
    GODE_LAUNCH_CHECK();
    return GODE_OK;
  }
  k_gemm_tc<<<grid, tc::WS_THREADS, smem, st>>>(M, N, K, A, lda, Bt, ldb, C, ldc, bias, relu,
                                                precision == GODE_PREC_TF32 ? 1 : 3, vec_in, vec_out, k_splits, band, pf_stages);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

}  // namespace gode

// C[M, N] = act(A[M, K] * Bt[N, K]^T + bias) on tcgen05 (3xTF32, or single-pass TF32 with GODE_PREC_TF32).
extern "C" int gode_gemm_tc_f32(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* Bt, int64_t ldb,
                                const float* bias, int32_t relu, float* C, int64_t ldc, int32_t precision, void* stream) {
  return gode::gemm_tc(M, N, K, A, lda, Bt, ldb, C, ldc, bias, relu, precision, gode::as_stream(stream), 1);
}

// The same product with the K range cut into k_splits slabs (reductions over millions of rows with a small M x N: weight
// gradients): slab s writes its partial product to C + s * M * ldc; the caller adds the k_splits partials in slab order.
extern "C" int gode_gemm_tc_splitk_f32(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* Bt, int64_t ldb,
                                       float* C_partials, int64_t ldc, int32_t k_splits, int32_t precision, void* stream) {
  return gode::gemm_tc(M, N, K, A, lda, Bt, ldb, C_partials, ldc, nullptr, 0, precision, gode::as_stream(stream), k_splits);
}

// out[d, ncols] = GroupNorm-normalised(y)^T * G, cs = column sums of G (see gn_wgrad_tc).
extern "C" size_t gode_gn_wgrad_workspace_bytes(int32_t d) { return d > 0 ? gode::gn_wgrad_ws_bytes(d) : 0; }

extern "C" int gode_gn_wgrad_f32(int64_t n, int32_t d, int32_t groups, float eps, const float* y, const float* G, int64_t ldg,
                                 int32_t ncols, float* out, int64_t ldo, float* cs, void* ws, size_t ws_bytes, int32_t precision,
                                 void* stream) {
  GODE_REQUIRE(n >= 0 && (n == 0 || (y && G)) && out && cs, "gn_wgrad: null pointer");
  return gode::gn_wgrad_tc(n, d, groups, eps, y, G, ldg, ncols, out, ldo, cs, static_cast<float*>(ws), ws_bytes, precision,
                           gode::as_stream(stream));
}

// out[n, ncols] = [t | GroupNorm(y)] * W  with W [d + 1, ncols] (row 0 the t row), see gn_linear_tc.
extern "C" int gode_gn_linear_supported(int32_t d, int32_t groups, int32_t ncols) {
  return gode::gn_linear_tc_supported(d, groups, ncols) ? 1 : 0;
}

extern "C" int gode_gn_linear_f32(int64_t n, int32_t d, int32_t groups, float eps, const float* y, const float* gamma,
                                  const float* beta, const float* W, int64_t ldw, float t, int32_t ncols, float* out, int64_t ldo,
                                  int32_t precision, void* stream) {
  GODE_REQUIRE(n >= 0 && (n == 0 || (y && out)) && gamma && beta && W, "gn_linear: null pointer");
  return gode::gn_linear_tc(n, d, groups, eps, y, gamma, beta, W, ldw, t, ncols, out, ldo, precision, gode::as_stream(stream));
}
