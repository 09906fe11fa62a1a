// libgode: the trainer's per-epoch epilogue -- log_softmax + masked nll_loss + accuracy in one pass, their
// backward in one pass, and Adam over a flat parameter buffer in one launch.
//
// Reference call sites: F.log_softmax(x, dim=1) at the end of every model (GCN/models.py:20,81,118,218 ...);
// F.nll_loss(output[idx_train], labels[idx_train]) and accuracy(output[idx_train], labels[idx_train])
// (GCN/train_res.py:76-77, GCN/utils.py:215-219); optim.Adam(lr, weight_decay) .step() (GCN/train_res.py:126-127, :79).
// At Cora scale an epoch is launch-bound (the whole state is 173 KB): these three kernels replace ~25 ATen launches
// (index_select x4, log_softmax, nll_loss, max, eq, sum, div, their backward, and six multi-tensor Adam kernels).
#include "internal.cuh"

namespace gode {

constexpr int LN_BLOCK = 256;   // 8 rows per block, one warp per row

// logp[r, :] = z[r, :] - logsumexp(z[r, :]);  for rows with mask[r] != 0:  loss_part += -logp[r, label[r]],
// correct_part += (argmax_c logp[r, c] == label[r])  (first maximum wins, as torch.max does)
__global__ void __launch_bounds__(LN_BLOCK) k_lsm_nll_fwd(int64_t n, int c, const float* __restrict__ z, int64_t ldz,
                                                          const int64_t* __restrict__ labels, const uint8_t* __restrict__ mask,
                                                          float* __restrict__ logp, int64_t ldo, float* __restrict__ part /*[grid][2]*/) {
  __shared__ float s_loss[LN_BLOCK / 32], s_ok[LN_BLOCK / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t r = blockIdx.x * (int64_t)(LN_BLOCK / 32) + w;
  float loss = 0.f, ok = 0.f;
  if (r < n) {
    const float* zr = z + r * ldz;
    float mx = -INFINITY;
    int arg = 0x7fffffff;
    for (int j = lane; j < c; j += 32) {
      const float v = __ldg(zr + j);
      if (v > mx) { mx = v; arg = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    float se = 0.f;
    for (int j = lane; j < c; j += 32) se += expf(__ldg(zr + j) - mx);
    se = warp_sum(se);
    const float lse = mx + logf(se);
    for (int j = lane; j < c; j += 32) logp[r * ldo + j] = __ldg(zr + j) - lse;
    if (lane == 0 && mask[r]) {
      const int64_t y = labels[r];
      loss = -(__ldg(zr + y) - lse);
      ok = (arg == static_cast<int>(y)) ? 1.f : 0.f;
    }
  }
  if (lane == 0) { s_loss[w] = loss; s_ok[w] = ok; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < LN_BLOCK / 32; ++i) { a += s_loss[i]; b += s_ok[i]; }
    part[2 * blockIdx.x] = a;
    part[2 * blockIdx.x + 1] = b;
  }
}

// out[0] = sum(part[:, 0]) / m,  out[1] = sum(part[:, 1]) / m   (one block, fixed order: deterministic)
__global__ void k_lsm_nll_finish(int nblk, const float* __restrict__ part, float inv_m, float* __restrict__ out) {
  __shared__ float s[2][32];
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) { a += part[2 * i]; b += part[2 * i + 1]; }
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = a; s[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float x = 0.f, y = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) { x += s[0][i]; y += s[1][i]; }
    out[0] = x * inv_m;
    out[1] = y * inv_m;
  }
}

// dz[r, j] = gloss * mask[r] / m * (exp(logp[r, j]) - [j == label[r]])
__global__ void k_lsm_nll_bwd(int64_t n, int c, const float* __restrict__ logp, int64_t ldo, const int64_t* __restrict__ labels,
                              const uint8_t* __restrict__ mask, const float* __restrict__ gloss, float inv_m,
                              float* __restrict__ dz, int64_t ldz) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * c) return;
  const int64_t r = i / c;
  const int j = static_cast<int>(i - r * c);
  float g = 0.f;
  if (mask[r]) g = (*gloss) * inv_m * (expf(logp[r * ldo + j]) - (labels[r] == j ? 1.f : 0.f));
  dz[r * ldz + j] = g;
}

// torch.optim.Adam (amsgrad off, maximize off): g += wd * p; m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2;
// p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void k_adam(int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       float lr, float b1, float b2, float eps, float wd, const int32_t* __restrict__ step /*device counter, incremented here*/) {
  const int t = *step + 1;
  const float bc1 = 1.f - powf(b1, static_cast<float>(t)), bc2 = 1.f - powf(b2, static_cast<float>(t));
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float pi = p[i];
    const float gi = g[i] + wd * pi;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}
__global__ void k_adam_tick(int32_t* step) { *step += 1; }

}  // namespace gode

using namespace gode;

extern "C" size_t gode_lsm_nll_workspace_bytes(int64_t n) {
  return align_up(sizeof(float) * 2 * static_cast<size_t>((n + LN_BLOCK / 32 - 1) / (LN_BLOCK / 32) + 1), 256);
}

extern "C" int gode_lsm_nll_fwd(int64_t n, int32_t c, const float* z, int64_t ldz, const int64_t* labels, const uint8_t* mask,
                                int64_t m, float* logp, int64_t ldo, float* loss_acc /*[2]*/, void* ws, size_t ws_bytes, void* stream) {
  GODE_REQUIRE(n >= 0 && c >= 1 && ldz >= c && ldo >= c && m >= 1, "lsm_nll_fwd: bad shape");
  GODE_REQUIRE(n == 0 || (z && labels && mask && logp && loss_acc && ws), "lsm_nll_fwd: null pointer");
  if (ws_bytes < gode_lsm_nll_workspace_bytes(n)) {
    set_error("lsm_nll_fwd: workspace too small");
    return GODE_EWORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  const int nblk = static_cast<int>((n + LN_BLOCK / 32 - 1) / (LN_BLOCK / 32));
  if (nblk > 0) {
    k_lsm_nll_fwd<<<nblk, LN_BLOCK, 0, st>>>(n, c, z, ldz, labels, mask, logp, ldo, static_cast<float*>(ws));
    GODE_LAUNCH_CHECK();
  }
  k_lsm_nll_finish<<<1, 256, 0, st>>>(nblk, static_cast<float*>(ws), 1.f / static_cast<float>(m), loss_acc);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

extern "C" int gode_lsm_nll_bwd(int64_t n, int32_t c, const float* logp, int64_t ldo, const int64_t* labels, const uint8_t* mask,
                                int64_t m, const float* gloss, float* dz, int64_t ldz, void* stream) {
  GODE_REQUIRE(n >= 0 && c >= 1 && ldz >= c && ldo >= c && m >= 1, "lsm_nll_bwd: bad shape");
  if (n == 0) return GODE_OK;
  GODE_REQUIRE(logp && labels && mask && gloss && dz, "lsm_nll_bwd: null pointer");
  const int64_t tot = n * c;
  k_lsm_nll_bwd<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, as_stream(stream)>>>(n, c, logp, ldo, labels, mask, gloss,
                                                                                        1.f / static_cast<float>(m), dz, ldz);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}

extern "C" int gode_adam_step(int64_t n, float* p, const float* g, float* m, float* v, float lr, float beta1, float beta2, float eps,
                              float weight_decay, int32_t* step_counter, void* stream) {
  GODE_REQUIRE(n >= 0 && (n == 0 || (p && g && m && v)) && step_counter, "adam_step: bad argument");
  cudaStream_t st = as_stream(stream);
  if (n > 0) {
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = 8LL * sm_count();
    if (blocks > cap) blocks = cap;
    k_adam<<<static_cast<unsigned>(blocks), 256, 0, st>>>(n, p, g, m, v, lr, beta1, beta2, eps, weight_decay, step_counter);
    GODE_LAUNCH_CHECK();
  }
  k_adam_tick<<<1, 1, 0, st>>>(step_counter);
  GODE_LAUNCH_CHECK();
  return GODE_OK;
}
