// L2 -> SM gather microbenchmark: how fast can this B200 deliver random 512-byte rows to the SMs
// when there is NO CSR (indices are hashed in registers), no epilogue and no heavy rows -- the ceiling the A_hat*S gather
// (k_spmm_vec, graph-odenet_b200/csrc/spmm.cu) is measured against.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/l2_gather_bench tools/l2_gather_bench.cu && /tmp/l2_gather_bench
//
// One warp per output row: DEG pseudo-random rows of a table of R rows (one float4 per lane = one coalesced 512 B request per
// row, exactly the access shape of the product kernel), U independent loads in flight per lane, summed, one 512 B row
// written per DEG rows read.  The table size sweeps from L2-resident (16 MB) to DRAM-resident (4 GB); "window" confines the
// rows of consecutive warps to a sliding band as a locality-ordered graph does.  Prints one JSON line per point with the
// bytes delivered to the SMs per second (rows * 512 B / time).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int U, int MINB>
__global__ void __launch_bounds__(256, MINB) k_gather(const float4* __restrict__ table, uint32_t rows, uint32_t window, int deg,
                                                       int64_t n_out, float4* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t w = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n_out) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // band start: proportional position of this output row in the table (window == rows: fully random)
  const uint32_t base = window >= rows ? 0u : (uint32_t)((double)w / (double)n_out * (double)(rows - window));
  for (int j = 0; j < deg; j += U) {
    float4 x[U];
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const uint32_t r = base + mix((uint32_t)w * 2654435761u + (uint32_t)(j + q) * 40503u + 17u) % window;
      x[q] = __ldg(table + (size_t)r * 32 + lane);
    }
#pragma unroll
    for (int q = 0; q < U; ++q) { acc.x += x[q].x; acc.y += x[q].y; acc.z += x[q].z; acc.w += x[q].w; }
  }
  __stcs(out + w * 32 + lane, acc);
}

template <int U, int MINB>
static float run(const float4* table, uint32_t rows, uint32_t window, int deg, int64_t n_out, float4* out, int iters) {
  const unsigned grid = (unsigned)((n_out + 7) / 8);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) k_gather<U, MINB><<<grid, 256>>>(table, rows, window, deg, n_out, out);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) k_gather<U, MINB><<<grid, 256>>>(table, rows, window, deg, n_out, out);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / iters;
}

int main() {
  const int deg = 20;
  const int64_t n_out = 4 * 1000 * 1000;   // 4 M output rows x 20 gathered rows = 41 GB delivered per launch
  const size_t max_rows = (size_t)8 << 20; // 4 GB table
  float4 *table, *out;
  CK(cudaMalloc(&table, max_rows * 512));
  CK(cudaMalloc(&out, (size_t)n_out * 512));
  CK(cudaMemset(table, 0, max_rows * 512));
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  struct P { uint32_t rows, window; const char* what; };
  const P pts[] = {{32768, 32768, "16 MB table, random (L2-resident)"},
                   {131072, 131072, "64 MB table, random (L2-resident)"},
                   {(uint32_t)max_rows, 65536, "4 GB table, 32 MB sliding band (locality-ordered graph)"},
                   {(uint32_t)max_rows, (uint32_t)max_rows, "4 GB table, random (DRAM-resident)"}};
  for (const P& p : pts) {
    const float a = run<4, 8>(table, p.rows, p.window, deg, n_out, out, 5);
    const float b = run<8, 4>(table, p.rows, p.window, deg, n_out, out, 5);
    const double bytes = (double)n_out * deg * 512.0;
    printf("{\"bench\": \"l2_gather\", \"what\": \"%s\", \"rows_gathered\": %lld, \"bytes_to_sm\": %.0f, "
           "\"ms_u4_8cta\": %.3f, \"ms_u8_4cta\": %.3f, \"TBps_u4_8cta\": %.2f, \"TBps_u8_4cta\": %.2f, \"max_sm_khz\": %d}\n",
           p.what, (long long)(n_out * deg), bytes, a, b, bytes / a / 1e9, bytes / b / 1e9, clk_khz);
    fflush(stdout);
  }
  return 0;
}
