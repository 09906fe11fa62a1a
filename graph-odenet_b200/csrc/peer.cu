// libgode: halo exchange over NVLink peer memory for the row-partitioned (multi-GPU) path.
//
// The reference is single-device.  A row-partitioned torch.spmm(adj, support) (GCN/layers.py:71)
// needs, on every rank, the rows of `support` that other ranks own.  Instead of pack -> NCCL all-to-all-v ->
// halo tail (gode_gather_rows + host-issued collective), ONE kernel reads the rows its peers reference from the
// local operand and stores them straight into the halo tails of the peers' operand buffers through peer
// mappings (cudaIpc handles, NVLink / NVSwitch), then the last CTA to finish publishes an epoch flag to every
// peer; the consumer waits for all peers' flags (gode_peer_wait) before it gathers.
//
// Memory: operand buffers live in one arena per rank allocated with gode_peer_alloc (plain cudaMalloc, so it can
// be exported); every rank opens its peers' arenas once.  Buffers have the same offsets on every rank (SPMD).
//
// Ordering: stores are followed by __threadfence_system() in every CTA, the last CTA (atomic counter) fences
// again and writes the flags with st.release.sys; the waiter reads them with ld.acquire.sys.  The wait has a
// wall-clock limit (globaltimer) and reports GODE_PEER_TIMEOUT in the group's status word instead of hanging.
#include "internal.cuh"
#include <string.h>

namespace gode {

struct PushArgs {
  int32_t world, rank;
  int32_t n_seg;                          // peers with rows to send, visited in rotated order (rank+1, rank+2, ...)
  int64_t vstart[GODE_MAX_PEERS + 1];     // virtual row range of segment j
  int64_t send_off[GODE_MAX_PEERS];       // offset of segment j in send_idx
  float* dst[GODE_MAX_PEERS];             // peer address of the first row of segment j
  uint32_t* flag[GODE_MAX_PEERS];         // peer p's flag slot for this rank (all world-1 peers, p != rank)
  int32_t n_flag;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One sub-warp of d4 lanes moves one row (d4 * 16 bytes, a contiguous NVLink write); rows are visited in a
// rotated peer order so that at any moment the ranks write to different destinations.
__global__ void __launch_bounds__(256) k_halo_push4(PushArgs a, const int32_t* __restrict__ send_idx, int d4,
                                                    const float4* __restrict__ src, int64_t lds4, int64_t ldd4,
                                                    unsigned int* __restrict__ counter, uint32_t epoch, int signal) {
  const int64_t n_rows = a.vstart[a.n_seg];
  const int64_t total = n_rows * d4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / d4;
    const int c = static_cast<int>(i - v * d4);
    int j = 0;
#pragma unroll 1
    while (j + 1 < a.n_seg && v >= a.vstart[j + 1]) ++j;
    const int64_t k = v - a.vstart[j];
    const int s = __ldg(send_idx + a.send_off[j] + k);
    const float4 val = __ldg(src + (int64_t)s * lds4 + c);
    reinterpret_cast<float4*>(a.dst[j])[k * ldd4 + c] = val;
  }
  if (!signal) return;   // a part of a pipelined exchange: the last part publishes the epoch (same stream, in order)
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(counter, 1u);
    if (prev == gridDim.x - 1) {
      *counter = 0;                         // re-armed for the next exchange (stream-ordered)
      __threadfence_system();
      for (int p = 0; p < a.n_flag; ++p) st_release_sys(a.flag[p], epoch);
    }
  }
}

// flags_local[p] >= epoch for every peer p (wrap-safe), or time out
__global__ void k_peer_wait(int world, int rank, const uint32_t* __restrict__ flags_local, uint32_t epoch,
                            unsigned long long timeout_ns, int32_t* __restrict__ status) {
  const int p = threadIdx.x;
  if (p >= world || p == rank) return;
  const unsigned long long t0 = globaltimer_ns();
  unsigned spins = 0;
  while (static_cast<int32_t>(ld_acquire_sys(flags_local + p) - epoch) < 0) {
    if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > timeout_ns) {
      // A peer never published this epoch.  The kernels queued behind this wait would gather from a stale halo tail and
      // every result after it would be silently wrong, so the time-out is FATAL for the stream: record it (readable through
      // gode_peer_status while the context lives), say which peer, and trap -- every later CUDA call of the process then
      // fails, which the host sees at its next synchronisation.
      atomicExch(status, GODE_PEER_TIMEOUT);
      __threadfence_system();
      printf("libgode: rank %d timed out after %llu ms waiting for peer %d to publish halo epoch %u -- aborting the stream\n",
             rank, timeout_ns / 1000000ull, p, epoch);
      __trap();
    }
  }
}

}  // namespace gode

using namespace gode;

extern "C" int gode_peer_alloc(size_t bytes, void** out) {
  GODE_REQUIRE(out != nullptr && bytes > 0, "peer_alloc: bad argument");
  void* p = nullptr;
  GODE_CHECK_CUDA(cudaMalloc(&p, bytes));
  const size_t head = bytes < GODE_PEER_HEADER_BYTES ? bytes : static_cast<size_t>(GODE_PEER_HEADER_BYTES);
  cudaError_t e = cudaMemset(p, 0, head);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("peer_alloc: %s", cudaGetErrorString(e));
    return GODE_ECUDA;
  }
  *out = p;
  return GODE_OK;
}

extern "C" int gode_peer_free(void* p) {
  if (p) GODE_CHECK_CUDA(cudaFree(p));
  return GODE_OK;
}

extern "C" int gode_peer_export(const void* p, void* handle64) {
  GODE_REQUIRE(p && handle64, "peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == GODE_PEER_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t h;
  GODE_CHECK_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(p)));
  memcpy(handle64, &h, sizeof(h));
  return GODE_OK;
}

extern "C" int gode_peer_open(const void* handle64, void** out) {
  GODE_REQUIRE(handle64 && out, "peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  GODE_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *out = p;
  return GODE_OK;
}

extern "C" int gode_peer_close(void* p) {
  if (p) GODE_CHECK_CUDA(cudaIpcCloseMemHandle(p));
  return GODE_OK;
}

static int halo_push_impl(const gode_peer_group_t* g, uint32_t epoch, const int32_t* send_idx, const int64_t* seg_begin,
                          const int64_t* seg_end, const int64_t* dst_row, int64_t buf_offset, int32_t d, const float* src,
                          int64_t lds, int64_t ldd, int32_t max_ctas, int32_t signal, void* stream) {
  GODE_REQUIRE(g && g->world >= 1 && g->world <= GODE_MAX_PEERS && g->rank >= 0 && g->rank < g->world,
               "halo_push: bad peer group");
  GODE_REQUIRE(seg_begin && seg_end && dst_row && d > 0 && d % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "halo_push: bad shape");
  GODE_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (buf_offset & 15) == 0, "halo_push: misaligned operand");
  if (g->world == 1) return GODE_OK;
  PushArgs a;
  memset(&a, 0, sizeof(a));
  a.world = g->world;
  a.rank = g->rank;
  int64_t v = 0;
  for (int s = 1; s < g->world; ++s) {
    const int p = (g->rank + s) % g->world;
    GODE_REQUIRE(g->base[p] != nullptr, "halo_push: peer arena not opened");
    a.flag[a.n_flag++] = reinterpret_cast<uint32_t*>(g->base[p]) + g->rank;
    const int64_t cnt = seg_end[p] - seg_begin[p];
    GODE_REQUIRE(cnt >= 0, "halo_push: negative segment");
    if (cnt == 0) continue;
    a.vstart[a.n_seg] = v;
    a.send_off[a.n_seg] = seg_begin[p];
    a.dst[a.n_seg] = reinterpret_cast<float*>(static_cast<char*>(g->base[p]) + buf_offset) + dst_row[p] * ldd;
    ++a.n_seg;
    v += cnt;
  }
  a.vstart[a.n_seg] = v;
  GODE_REQUIRE(v == 0 || (send_idx && src), "halo_push: null pointer");
  if (v == 0 && !signal) return GODE_OK;
  const int d4 = d / 4;
  int64_t blocks = (v * d4 + 255) / 256;
  int64_t cap = max_ctas > 0 ? max_ctas : 8LL * sm_count();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  unsigned int* counter = reinterpret_cast<unsigned int*>(static_cast<char*>(g->base[g->rank]) + GODE_PEER_COUNTER_OFFSET);
  k_halo_push4<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(
      a, send_idx, d4, reinterpret_cast<const float4*>(src), lds / 4, ldd / 4, counter, epoch, signal);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

extern "C" int gode_halo_push(const gode_peer_group_t* g, uint32_t epoch, const int32_t* send_idx,
                              const int64_t* send_ptr, const int64_t* dst_row, int64_t buf_offset, int32_t d,
                              const float* src, int64_t lds, int64_t ldd, int32_t max_ctas, void* stream) {
  GODE_REQUIRE(send_ptr != nullptr, "halo_push: null send_ptr");
  return halo_push_impl(g, epoch, send_idx, send_ptr, send_ptr + 1, dst_row, buf_offset, d, src, lds, ldd, max_ctas, 1, stream);
}

extern "C" int gode_halo_push_part(const gode_peer_group_t* g, uint32_t epoch, const int32_t* send_idx,
                                   const int64_t* seg_begin, const int64_t* seg_end, const int64_t* dst_row,
                                   int64_t buf_offset, int32_t d, const float* src, int64_t lds, int64_t ldd,
                                   int32_t max_ctas, int32_t signal, void* stream) {
  return halo_push_impl(g, epoch, send_idx, seg_begin, seg_end, dst_row, buf_offset, d, src, lds, ldd, max_ctas, signal, stream);
}

extern "C" int gode_peer_wait(const gode_peer_group_t* g, uint32_t epoch, uint64_t timeout_ns, void* stream) {
  GODE_REQUIRE(g && g->world >= 1 && g->world <= GODE_MAX_PEERS && g->rank >= 0 && g->rank < g->world,
               "peer_wait: bad peer group");
  if (g->world == 1) return GODE_OK;
  char* base = static_cast<char*>(g->base[g->rank]);
  GODE_REQUIRE(base != nullptr, "peer_wait: local arena missing");
  k_peer_wait<<<1, 32, 0, as_stream(stream)>>>(g->world, g->rank, reinterpret_cast<const uint32_t*>(base), epoch,
                                               timeout_ns, reinterpret_cast<int32_t*>(base + GODE_PEER_STATUS_OFFSET));
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

extern "C" int gode_peer_status(const gode_peer_group_t* g, int32_t* status_host, void* stream) {
  GODE_REQUIRE(g && status_host && g->rank >= 0 && g->rank < GODE_MAX_PEERS && g->base[g->rank], "peer_status: bad argument");
  char* base = static_cast<char*>(g->base[g->rank]);
  GODE_CHECK_CUDA(cudaMemcpyAsync(status_host, base + GODE_PEER_STATUS_OFFSET, sizeof(int32_t), cudaMemcpyDeviceToHost,
                                  as_stream(stream)));
  GODE_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return GODE_OK;
}
