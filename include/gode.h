/*
 * gode.h -- C ABI of libgode.so, the B200 (sm_100a) hot path of graph-odenet.
 *
 * The reference (phcavelar/graph-odenet) is pure Python: its "operator interface" for this path is the
 * set of ATen calls its layers make.  Each entry point below names the reference call site it replaces
 * (file:line under /root/reference).  The reference-side binding (a ctypes stub) is in INTEGRATION.md.
 *
 * Conventions
 *   - every function is extern "C", takes plain pointers and sizes, returns 0 on success and a negative
 *     GODE_E* code on failure; gode_last_error() returns a thread-local message for the last failure;
 *   - all pointers are DEVICE pointers unless the name ends in _host; all matrices are row-major fp32
 *     with an explicit leading dimension where one is given; indices are int32 (the plan builder takes
 *     the reference's int64 COO and range-checks it);
 *   - nothing here allocates caller-visible memory: workspaces are sized by *_workspace_bytes() and
 *     owned by the caller (the PyTorch caching allocator in the shipped host code);
 *   - every call only enqueues work on `stream` (a cudaStream_t passed as void*); none synchronises;
 *   - there is no CPU fallback: without a CUDA device every compute call fails with GODE_ENODEV.
 */
#ifndef GODE_H_
#define GODE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GODE_OK 0
#define GODE_EINVAL (-1)   /* bad argument (shape, alignment, null pointer) */
#define GODE_ECUDA (-2)    /* a CUDA runtime call failed; see gode_last_error() */
#define GODE_ENODEV (-3)   /* no usable CUDA device */
#define GODE_EWORKSPACE (-4) /* workspace too small */

#define GODE_MAX_STAGES 8
#define GODE_MAX_PEERS 16   /* ranks of one NVLink domain the peer-memory halo exchange addresses */

/* precision of the dense H*W products (GCN/layers.py:32,70 torch.mm) */
#define GODE_PREC_FP32 0   /* fp32 result: SIMT FFMA, or 3xTF32 split on tcgen05 where the shape allows */
#define GODE_PREC_TF32 1   /* single-pass TF32 on tcgen05 (flagged option) */

int gode_version(void);
const char* gode_last_error(void);
/* sm_count / compute capability of the current device */
int gode_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* number of CUDA kernels this library has launched in this process (monotonic) */
unsigned long long gode_launch_count(void);

/* Leave n SMs free of the library's persistent (one-CTA-per-SM) kernels, so that a collective issued on another
 * stream (the NCCL halo exchange of the row-partitioned path) can start underneath them instead of queueing behind
 * them: a persistent tcgen05 kernel holds ~200 KB of shared memory on every SM for its whole duration.  Default 0. */
int gode_reserve_sms(int n);

/* Optional CUDA-event timing of the library's kernel classes, recorded on the launching stream
 * (used by bench.py for the roofline of the dominant kernel; off by default, ~2 us per scope when on).
 * enable(on) clears the records.  read() synchronises on the recorded events. */
#define GODE_PROF_AGG_FWD 0    /* A_hat * S gather with fused epilogue (the ODE function's SpMM)   */
#define GODE_PROF_AGG_T 1      /* A_hat^T * gP gather of the VJP                                     */
#define GODE_PROF_TRANSFORM 2  /* S = [t || GN(y)] W                                                 */
#define GODE_PROF_VJP_DENSE 3  /* column sums, weight-gradient and input-gradient products, GN bwd   */
#define GODE_PROF_OTHER 4      /* stand-alone gode_spmm / gode_gemm / rk helpers                     */
#define GODE_PROF_KINDS 5
int gode_profile_enable(int on);
int gode_profile_read(int kind, int* launches, float* total_ms, float* max_ms);

/* ------------------------------------------------------------------------------------------------
 * Graph plan: canonical CSR from the reference's COO.
 * replaces: the implicit COO->CSR conversion inside torch.spmm (GCN/layers.py:33,71) fed by
 *           sparse_mx_to_torch_sparse_tensor (GCN/utils.py:222-229).
 * Entries are ordered by (row, col), stable; entries with equal (row, col) are summed in fp32 in input
 * order.  *nnz_out (device int64) receives the number of stored entries after merging.
 * rowptr [n_rows+1], colidx/vals [nnz] must be provided at full nnz capacity.
 * ---------------------------------------------------------------------------------------------- */
size_t gode_csr_from_coo_workspace_bytes(int64_t nnz, int64_t n_rows);
int gode_csr_from_coo(int64_t n_rows, int64_t n_cols, int64_t nnz,
                      const int64_t* row, const int64_t* col, const float* val,
                      int32_t* rowptr, int32_t* colidx, float* vals, int64_t* nnz_out,
                      void* ws, size_t ws_bytes, void* stream);

/* CSR of the transpose (entries ordered by (col,row)); perm_t[e'] = index of the source entry.
 * Needed because the reference's row-normalised adjacency is not symmetric (GCN/utils.py:186,205-212):
 * autograd's backward of torch.spmm multiplies by A^T. */
size_t gode_csr_transpose_workspace_bytes(int64_t nnz, int64_t n_rows, int64_t n_cols);
int gode_csr_transpose(int64_t n_rows, int64_t n_cols, int64_t nnz,
                       const int32_t* rowptr, const int32_t* colidx, const float* vals,
                       int32_t* rowptr_t, int32_t* colidx_t, float* vals_t, int32_t* perm_t,
                       void* ws, size_t ws_bytes, void* stream);

/* A CSR matrix as the kernels consume it.  Rows with more than GODE_HEAVY_ROW stored entries ("heavy" rows:
 * the hubs of a power-law graph) are listed separately: the SpMM splits each of them into chunks of
 * GODE_HEAVY_CHUNK entries that are gathered by independent warps and summed in chunk order, so one hub
 * cannot serialise the kernel.  heavy_chunk_ptr[i] = first chunk of heavy row i (exclusive prefix sum of
 * ceil(len/GODE_HEAVY_CHUNK)), heavy_chunk_ptr[n_heavy] = n_chunks. */
#define GODE_HEAVY_ROW 256
#define GODE_HEAVY_CHUNK 256
typedef struct {
  int64_t n_rows;
  int64_t n_cols;
  const int32_t* rowptr;            /* [n_rows + 1] */
  const int32_t* colidx;            /* [nnz] */
  const float* vals;                /* [nnz] */
  const float* row_vals;            /* [n_rows] or NULL.  When every stored entry of a row carries the same value
                                       (the reference's D^-1 (A+I), GCN/utils.py:186,205-212: 1/(deg+1)), the SpMM
                                       gathers with colidx only and scales the row sum once (gode_csr_row_values). */
  const int32_t* heavy_rows;        /* [n_heavy] ascending, or NULL */
  const int32_t* heavy_chunk_ptr;   /* [n_heavy + 1], or NULL */
  int32_t n_heavy;
  int32_t n_chunks;
  const int32_t* tile_sched;        /* [n_tile_sched] or NULL.  Work order of the d = 128 tile gather (one CTA per entry): an entry
                                       >= 0 is a 32-row tile index; an entry < 0 is -(g + 1), the hub chunks [8g, 8g + 8).  Tiles are
                                       in id order and every chunk group follows the tile that contains its (first chunk's) hub, so
                                       a hub's chunks are gathered while the band around its id is L2-resident; NULL: the hub chunks
                                       are gathered by a separate launch after the tiles (same sums, same order, same result). */
  int32_t n_tile_sched;
} gode_csr_t;

/* row_vals_out[r] = the common value of row r's entries (0 for an empty row); *is_const_out (device int32) = 1
 * iff every row's entries are bit-identical -- only then may row_vals_out be used as gode_csr_t.row_vals. */
int gode_csr_row_values(int64_t n_rows, const int32_t* rowptr, const float* vals, float* row_vals_out,
                        int32_t* is_const_out, void* stream);

/* heavy_rows (capacity n_rows) and heavy_chunk_ptr (capacity n_rows + 1) are filled on the device;
 * counts_out is a device int32[2] = {n_heavy, n_chunks}. */
int gode_csr_heavy_rows(int64_t n_rows, const int32_t* rowptr, int32_t* heavy_rows, int32_t* heavy_chunk_ptr,
                        int32_t* counts_out, void* stream);

/* Route of a fused halo push (multi-GPU, peer memory; see the "Halo exchange over NVLink peer memory" section):
 * the kernel that PRODUCES a gather operand stores every row a peer references straight into the halo tail of that
 * peer's operand buffer, next to its own local store, so the transfer rides underneath the producing kernel.
 * ptr == NULL: no push.  Entries of local row r: ent[ptr[r] .. ptr[r+1]) = (peer << 40) | destination row. */
typedef struct {
  const int32_t* ptr;             /* device, [n_rows + 1] */
  const int64_t* ent;             /* device */
  float* base[GODE_MAX_PEERS];    /* peer p's operand buffer through the peer mapping (same leading dimension) */
} gode_push_route_t;

/* ------------------------------------------------------------------------------------------------
 * CSR SpMM with a fused row epilogue.
 * replaces: torch.spmm(adj, support) (+ bias, + F.relu, + residual) -- GCN/layers.py:33-35,71-73,
 *           GCN/models.py:76-80,110-116,178.
 *   acc   = sum_e vals[e] * X[colidx[e], :]   (+ acc_in, if given)
 *   v     = acc + bias (if bias)          ; v = max(v,0) (if relu)
 *   Y     = v + residual (if residual)    (Y may be NULL)
 *   fused Runge-Kutta stage combination (torchdiffeq rk_common._runge_kutta_step, restated in
 *   oracle/odeint.py):  Ynext = y0 + sum_j coef[j]*kprev[j] + coef_self*v     (if ynext)
 *   fused adjoint preparation:   gP = mask_scale * mask_src * (v > 0)         (if gp_out)
 *   second combination of the SAME operands (if second.out; needs y0):
 *                                out = y0 + sum_j second.coef[j]*kprev[j] + second.coef_self*v
 *     -- the solver keeps a running partial sum of the step's final weights b this way (written INSTEAD of the
 *        stage derivative), so that the last stage of a step reads one tensor rather than every earlier k_j.
 *   gp_row_scale (with gp_out): gP rows are multiplied by gp_row_scale[row] -- the A_hat^T gather of a row-constant
 *     A_hat then needs no values stream (column i of A_hat^T is the constant row value of row i of A_hat).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  float coef[GODE_MAX_STAGES];
  float coef_self;
  float* out;              /* [n_rows, ld] or NULL: off */
} gode_rk_second_t;

typedef struct {
  const float* bias;       /* [d] or NULL */
  int32_t relu;            /* 0/1 */
  const float* residual;   /* [n_rows, ld] or NULL */
  const float* y0;         /* RK base state or NULL */
  const float* kprev[GODE_MAX_STAGES];
  float coef[GODE_MAX_STAGES];
  int32_t n_prev;
  float coef_self;
  float* ynext;            /* [n_rows, ld] or NULL */
  const float* mask_src;   /* adjoint state a, or NULL */
  float mask_scale;
  float* gp_out;           /* [n_rows, ld] or NULL */
  const float* acc_in;     /* [n_rows, ld] or NULL: partial sums added to acc before the bias (the product with
                              another column block of the same rows -- the row-partitioned path gathers the owned
                              columns while the halo is in flight, then the halo columns with acc_in) */
  gode_push_route_t push;  /* gp_out rows are also stored into the peers' buffers (ptr NULL: off) */
  gode_push_route_t push_y;/* ynext rows are also stored into the peers' buffers (ptr NULL: off): the row-partitioned solver
                              exchanges the transform's INPUT this way and transforms [owned | halo] rows locally, so no
                              exchange of the support S stands between two gathers */
  gode_rk_second_t second; /* second combination of (y0, kprev, v); out NULL: off */
  const float* gp_row_scale; /* [n_rows] or NULL */
} gode_spmm_epilogue_t;

/* ws: n_chunks * d floats of scratch for the heavy-row partial sums (0 bytes when n_heavy == 0) */
size_t gode_spmm_workspace_bytes(const gode_csr_t* A, int32_t d);
int gode_spmm_csr_f32(const gode_csr_t* A, const float* X, int64_t ldx, int32_t d, float* Y, int64_t ldy,
                      const gode_spmm_epilogue_t* epi, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense GEMM  C = alpha * op(A) op(B) + beta * C      (row-major, fp32 in/out)
 * replaces: torch.mm(input, weight) -- GCN/layers.py:32,70; its autograd transposes.
 * splits > 1 partitions K (used for the tall weight-gradient products); ws must hold
 * splits*M*N floats; the partial sums are reduced in a fixed order (deterministic).
 * ---------------------------------------------------------------------------------------------- */
int gode_gemm_f32(int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K, float alpha,
                  const float* A, int64_t lda, const float* B, int64_t ldb, float beta,
                  float* C, int64_t ldc, int32_t precision, int32_t splits, void* ws, size_t ws_bytes,
                  void* stream);

/* C = act(A op(B) + bias)  -- a Linear layer with bias (and ReLU) in the GEMM epilogue.
 * replaces: nn.Linear f(h), w(h) (GAT/layers.py:20-21,44-45, as node projections: see gode_gat_fwd);
 *           MyLinear / MLP (QC/layers.py:26-30, 49-57: mm + bias, relu between layers).
 * transB = 0: B is [K, N] (MyLinear's weight layout); transB = 1: B is [N, K] (nn.Linear's). bias [N] or NULL. */
int gode_linear_f32(int32_t transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                    const float* B, int64_t ldb, const float* bias, int32_t relu, float* C, int64_t ldc,
                    void* stream);

/* The same Linear on the tensor cores (tcgen05, fp32 result by 3xTF32; precision = GODE_PREC_TF32 for one pass):
 *   C[M, N] = act(A[M, K] * Bt[N, K]^T + bias)      -- both operands K-major: a weight stored [K, N] is transposed first.
 * Arbitrary M, N, K (edges are zero-filled / masked); 128-bit accesses when lda, ldb, ldc are multiples of 4 and the bases
 * 16-byte aligned.  Replaces torch.mm / nn.Linear at QC/layers.py:76-86 (the edge encoder, [E, 2667] x [2667, 5329]),
 * GAT/layers.py:40-45 (the f / w projections) and GCN/layers.py:32 (input layer).  Accumulation is hierarchical (K slabs of
 * 256 in TMEM, fp32 running sums outside): as accurate as an fp32 SGEMM for any K (tests/test_gpu_gemm_tc.py). */
int gode_gemm_tc_f32(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* Bt, int64_t ldb,
                     const float* bias, int32_t relu, float* C, int64_t ldc, int32_t precision, void* stream);

/* The same product with the K range cut into k_splits slabs -- for reductions over very many rows with a small M x N
 * (weight gradients x^T g: torch.autograd's mm at the call sites above).  Slab s writes its partial product to
 * C_partials + s * M * ldc; the caller adds the k_splits partials in slab order (deterministic). */
int gode_gemm_tc_splitk_f32(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* Bt, int64_t ldb,
                            float* C_partials, int64_t ldc, int32_t k_splits, int32_t precision, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Trainer epilogue: replaces F.log_softmax (GCN/models.py:20,81,118,218), F.nll_loss(output[idx], labels[idx])
 * and accuracy(...) (GCN/train_res.py:76-77, GCN/utils.py:215-219), their autograd, and optim.Adam.step()
 * (GCN/train_res.py:79,126-127).  mask[r] != 0 marks the m rows of the index set; labels are int64 class ids per row.
 * loss_acc[0] = -mean_{masked r} logp[r, labels[r]], loss_acc[1] = mean_{masked r} (argmax logp[r] == labels[r]). */
size_t gode_lsm_nll_workspace_bytes(int64_t n);
int gode_lsm_nll_fwd(int64_t n, int32_t c, const float* z, int64_t ldz, const int64_t* labels, const uint8_t* mask, int64_t m,
                     float* logp, int64_t ldo, float* loss_acc, void* ws, size_t ws_bytes, void* stream);
/* dz = gloss[0] * d loss / d z   (gloss: device scalar, the upstream gradient of the loss) */
int gode_lsm_nll_bwd(int64_t n, int32_t c, const float* logp, int64_t ldo, const int64_t* labels, const uint8_t* mask, int64_t m,
                     const float* gloss, float* dz, int64_t ldz, void* stream);
/* torch.optim.Adam's update (amsgrad / maximize off, L2 weight decay added to the gradient) over flat fp32 buffers;
 * step_counter is a device int32 holding the number of steps taken so far, incremented by the call (so the call can sit
 * inside a replayed CUDA graph). */
int gode_adam_step(int64_t n, float* p, const float* g, float* m, float* v, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int32_t* step_counter, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row-wise GroupNorm on [n, d] (groups of d/groups contiguous channels per row).
 * replaces: nn.GroupNorm(min(32,d), d) -- GCN/models.py:165,175 (ATen formula
 *           y = x*(gamma*rstd) + (beta - mean*gamma*rstd)).
 * bwd: dx, and column reductions dgamma/dbeta written (not accumulated) to [d] each.
 * ws for bwd: gode_colreduce_workspace_bytes(2*d).
 * ---------------------------------------------------------------------------------------------- */
int gode_groupnorm_fwd(int64_t n, int32_t d, int32_t groups, float eps, const float* x, int64_t ldx,
                       const float* gamma, const float* beta, float* y, int64_t ldy, void* stream);
int gode_groupnorm_bwd(int64_t n, int32_t d, int32_t groups, float eps, const float* x, int64_t ldx,
                       const float* gamma, const float* dy, int64_t lddy, float* dx, int64_t lddx,
                       float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream);

/* out[d, ncols] (row stride ldo) = xhat(y)^T * G and cs[ncols] = column sums of G, with xhat = GroupNorm(y) BEFORE its
 * affine, reduced over the n rows on tcgen05 (3xTF32, hierarchical accumulation).  The weight gradient of a layer fed by
 * [t | GroupNorm(y)] whose output is wider than d -- the GAT ODE function's projection (GAT/models.py:161-179 through
 * GAT/layers.py:40-45): dW[1:] = gamma * out + beta * cs, dW[0] = t * cs, db = cs.  d = 128 with 32 groups; ncols and ldg
 * multiples of 4, 16-byte aligned operands; anything else returns GODE_EINVAL (callers use gode_gemm_f32 then). */
/* out[n, ncols] (row stride ldo) = [t*1 | GroupNorm(y)] * W, W [d+1, ncols] row-major (row stride ldw, row 0 the t row):
 * the node projection of the GAT ODE function (GAT/models.py:172-179 feeding GAT/layers.py:40-45) without the [N, d+1]
 * concatenation; GroupNorm is applied inside the producer of the tcgen05 product.  gode_gn_linear_supported: d = 128 with
 * 32 groups, ncols = k*128 + {0, 4, 8, 16, 32}. */
int gode_gn_linear_supported(int32_t d, int32_t groups, int32_t ncols);
int gode_gn_linear_f32(int64_t n, int32_t d, int32_t groups, float eps, const float* y, const float* gamma, const float* beta,
                       const float* W, int64_t ldw, float t, int32_t ncols, float* out, int64_t ldo, int32_t precision,
                       void* stream);

size_t gode_gn_wgrad_workspace_bytes(int32_t d);
int gode_gn_wgrad_f32(int64_t n, int32_t d, int32_t groups, float eps, const float* y, const float* G, int64_t ldg,
                      int32_t ncols, float* out, int64_t ldo, float* cs, void* ws, size_t ws_bytes, int32_t precision,
                      void* stream);

size_t gode_colreduce_workspace_bytes(int32_t width);
/* out[c] = sum_r x[r, c]   (bias gradients; GCN/layers.py:35 autograd) */
int gode_colsum_f32(int64_t n, int32_t d, const float* x, int64_t ldx, float* out,
                    void* ws, size_t ws_bytes, void* stream);

/* res = g * (out > 0): backward of a ReLU fused into a producing kernel's epilogue
 * (autograd of F.relu, GCN/models.py:76-80; MLP hidden layers QC/layers.py:55). */
int gode_relu_bwd(int64_t n_elems, const float* g, const float* out, float* res, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Runge-Kutta helpers (torchdiffeq rk_common / dopri5, restated in oracle/odeint.py).
 * combine : out = y0 + sum_j coef[j]*k[j]                 (y0 may be NULL -> plain linear combination)
 * error   : sumsq_out[0] = sum_i ( (sum_j coef[j]*k[j][i]) / (atol + rtol*max(|y0_i|,|y1_i|)) )^2
 *           (the caller divides by n: "mean-square error ratio", accept iff <= 1)
 * ws for error: gode_colreduce_workspace_bytes(1).
 * ---------------------------------------------------------------------------------------------- */
int gode_rk_combine(int64_t n_elems, const float* y0, const float* const* k_host, const float* coef_host,
                    int32_t n_k, float* out, void* stream);
int gode_rk_error_sumsq(int64_t n_elems, const float* y0, const float* y1, const float* const* k_host,
                        const float* coef_host, int32_t n_k, float rtol, float atol, float* sumsq_out,
                        void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GCN ODE function   f(t, y) = relu( A_hat * ([t || GroupNorm(y)] * W) + b )
 * replaces: ODEfunc.forward (GCN/models.py:172-179) = GroupNorm + cat + FixedGraphConvolution.forward
 *           (GCN/layers.py:69-75) + F.relu; and, for the adjoint, the torch.autograd.grad call
 *           torchdiffeq makes on it (restated in oracle/odeint.py:_Adjoint).
 *
 * The function is evaluated in two halves so the Runge-Kutta stage combination can be fused between
 * them ("support" S is the only [N,d] tensor that crosses a stage boundary):
 *   transform : S = [t || GroupNorm(y)] * W                               (dense, row-local)
 *   aggregate : k = relu(A_hat * S + b), y_next = y0 + sum c_j k_j, S_next = transform(y_next, t_next)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  gode_csr_t A;            /* A_hat row block: n_rows = rows owned (outputs), n_cols = rows of the gather operand S
                              (= n_rows, or owned + halo when partitioned) */
  gode_csr_t At;           /* A_hat^T row block (rows owned; n_cols = owned + halo of the transpose); rowptr may be
                              NULL when only forward evaluations are needed */
  int32_t d;
  int32_t groups;
  float gn_eps;
  int32_t precision;       /* GODE_PREC_* */
  const float* W;          /* [d+1, d]; row 0 multiplies the time column */
  const float* b;          /* [d] or NULL */
  const float* gamma;      /* [d] */
  const float* beta;       /* [d] */
  /* Split gathers of the row-partitioned path (both 0 / NULL otherwise): the gather operand of A (S) and of At (gP)
   * is read from row `gather_row_offset` on (column indices are relative to it) and `partial_in` [n_rows, d] is
   * added to the gathered sums before the epilogue.  Everything else (transform output, column sums of gP, ...)
   * still addresses the buffers from row 0. */
  int64_t gather_row_offset;
  const float* partial_in;
  /* Fused halo pushes (ptr NULL: off): rows of every support S the transform writes / of every gP that phase 1 of the
   * VJP writes are also stored into the peers' halo tails.  Needs the tensor-core transform (gode_gcn_push_fusable). */
  gode_push_route_t push_S;
  gode_push_route_t push_gP;
  gode_push_route_t push_y;   /* rows of every y_next the fused Runge-Kutta combination writes (gode_gcn_stage_fwd, _stage_fwd_rows,
                                 _vjp_phase1); the caller then transforms owned AND halo rows (gode_gcn_transform_rows) */
  /* Per-call (out NULL: off): the second Runge-Kutta combination of the NEXT gode_gcn_stage_fwd / _stage_fwd_rows /
   * _vjp_phase1 (over y0, kprev, k) or gode_gcn_vjp_phase2_rk (over a0, kprev, k_a) call -- see gode_spmm_epilogue_t. */
  gode_rk_second_t second;
  /* Non-NULL ([n_rows], = A.row_vals of the FULL row block): A_hat is row-constant AND row-stochastic (the reference's
   * D^-1 (A + I), GCN/utils.py:205-212) and At holds the 0/1 pattern of A_hat^T (values and row_vals all one).  Phase 1
   * then stores gP pre-multiplied by gp_row_scale[row] (column i of A_hat^T is the constant row value of row i), the
   * A_hat^T gather reads no values, and the bias gradient (column sums of the unscaled gP) is taken from the column sums
   * of gS = A_hat^T gP, which the weight-gradient pass forms anyway: sum_i (sum_j A_ij) gP_i = sum_i gP_i.
   * NULL: values per entry, separate column-sum pass over gP. */
  const float* gp_row_scale;
} gode_gcn_odefunc_t;

size_t gode_gcn_workspace_bytes(const gode_gcn_odefunc_t* f);
/* 1 when the kernels this descriptor selects can carry push_S / push_gP, else 0 */
int gode_gcn_push_fusable(const gode_gcn_odefunc_t* f);

/* S[n_rows, d] = [t || GN(y)] W */
int gode_gcn_transform(const gode_gcn_odefunc_t* f, const float* y, float t, float* S,
                       void* ws, size_t ws_bytes, void* stream);
/* The same transform on rows [row0, row0 + n_rows) of the block (y and S are the block's base pointers): lets the caller
 * pipeline the transform of one row chunk with the halo push of the previous one.  With the tensor-core transform the range
 * may extend over the halo rows of a partitioned block (row0 + n_rows <= A.n_cols = owned + halo: y and S then have that
 * many rows) -- the transform is row-local, so a rank that holds its peers' rows of y (push_y) needs no exchange of S. */
int gode_gcn_transform_rows(const gode_gcn_odefunc_t* f, const float* y, float t, float* S, int64_t row0, int64_t n_rows,
                            void* ws, size_t ws_bytes, void* stream);

/* k_out = relu(A_hat S + b); optional y_next (RK combination, see gode_spmm_epilogue_t);
 * optional S_next = transform(y_next, t_next) (needs y_next). */
int gode_gcn_stage_fwd(const gode_gcn_odefunc_t* f, const float* S, float* k_out,
                       const float* y0, const float* const* kprev_host, const float* coef_host, int32_t n_prev,
                       float coef_self, float* y_next, float t_next, float* S_next,
                       void* ws, size_t ws_bytes, void* stream);

/* The same stage for the rows [row0, row0 + n_rows) of the block only (d = 128; no fused transform).  Lets a row-partitioned
 * caller pipeline a stage: chunk c's rows of y_next are transformed (gode_gcn_transform_rows) and pushed to the peers while
 * chunk c+1 is still being gathered. */
int gode_gcn_stage_fwd_rows(const gode_gcn_odefunc_t* f, const float* S, float* k_out, const float* y0,
                            const float* const* kprev, const float* coef, int32_t n_prev, float coef_self, float* y_next,
                            int64_t row0, int64_t n_rows, void* ws, size_t ws_bytes, void* stream);

/* One evaluation of the augmented (adjoint) dynamics at (t, y, a) with upstream = sign * a:
 *   k_y  = f(t, y)                                   [n_rows, d]
 *   k_a  = sign * a^T df/dy                          [n_rows, d]
 *   gtheta = sign * a^T df/d(theta, t), laid out as  [ W (d+1)*d | b d | gamma d | beta d | t 1 ]
 * S must hold transform(y, t).  (torchdiffeq passes -a: sign = -1.) */
int gode_gcn_stage_vjp(const gode_gcn_odefunc_t* f, const float* y, float t, const float* S,
                       const float* a, float sign, float* k_y, float* k_a, float* gtheta,
                       void* ws, size_t ws_bytes, void* stream);

/* The same evaluation in two halves, for the row-partitioned (multi-GPU) path where the rows of gP that
 * other ranks own must be exchanged between them:
 *   phase1: k_y = relu(A_hat S + b);  gP[0:n_rows] = sign * a * (k_y > 0);  optional fused RK combination
 *           y_next = y0 + sum_j coef[j] kprev[j] + coef_self k_y (the y-part of the augmented stage state)
 *   phase2: gS = A_hat^T gP (gP has n_cols_t rows);  k_a, gtheta as above (gtheta = this rank's partial sum) */
int gode_gcn_vjp_phase1(const gode_gcn_odefunc_t* f, const float* S, const float* a, float sign,
                        float* k_y, float* gP,
                        const float* y0, const float* const* kprev_host, const float* coef_host, int32_t n_prev,
                        float coef_self, float* y_next, void* ws, size_t ws_bytes, void* stream);
int gode_gcn_vjp_phase2(const gode_gcn_odefunc_t* f, const float* y, float t, const float* gP,
                        float* k_a, float* gtheta, void* ws, size_t ws_bytes, void* stream);
/* phase2 with the Runge-Kutta combination of the adjoint state fused behind the GroupNorm backward:
 *   a_next = a0 + sum_j coef[j] kprev[j] + coef_self k_a      (a_next may be null: plain phase2)
 * k_a may be null when no later stage reads it (then only a_next is written). */
int gode_gcn_vjp_phase2_rk(const gode_gcn_odefunc_t* f, const float* y, float t, const float* gP,
                           float* k_a, float* gtheta, const float* a0, const float* const* kprev_host,
                           const float* coef_host, int32_t n_prev, float coef_self, float* a_next,
                           void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GAT attention aggregation (segmented softmax / scatter over the edge list).
 * replaces: GAT/layers.py:40-58 and :103-120 -- x[src], x[tgt], cat, relu(f(h)), w(h), torch.max(a, 0), exp,
 *           the two torch.spmm(Mtgt, .) incidence products and the division.
 * The caller projects the nodes once (gode_gemm_f32):  P[N, ldp] = [Ps (C) | Pt (C) | as (H) | at (H)],
 *   Ps = x Wf[:, :i]^T, Pt = x Wf[:, i:]^T + bf, as = x ww[:, :i]^T, at = x ww[:, i:]^T + bw   (C = heads*oh),
 * so that z_e = Ps[src_e] + Pt[tgt_e] = f(h_e) and a_e = as[src_e] + at[tgt_e] = w(h_e).
 *   out[n, c] = sum_{e: tgt_e = n} relu(z_e[c]) exp(a_e - amax_h) / (sum_{e: tgt_e = n} exp(a_e - amax_h) + eps)
 * with amax_h the maximum of a_e over ALL edges (the reference's global shift), per head h = c / oh.
 * heads > 1 is the builder extension "H independent reference heads, concatenated" (SURVEY 8a).
 * Edges are grouped by target (tptr / t_src / t_tgt: CSR of Mtgt, edge order kept inside a segment) and by
 * source (sptr / s_tgt / s_pos; s_pos = position of the edge in the by-target order) -- built once per graph.
 * fwd saves den[N, heads] (denominators incl. eps) and amax_key[heads] (max value and arg-max edge, packed).
 * nan_flag (device int32) is set to 1 if any a_e or output is NaN (the reference asserts on those,
 * GAT/layers.py:46-56); it is never cleared by the library.
 * bwd writes dP[N, ldp] (columns [0, 2C+2H) are written, padding is left alone); the caller finishes with
 * dx = dP W^T and dW = x^T dP.
 * ---------------------------------------------------------------------------------------------- */
/* Hub nodes of one grouping: a node with more than GODE_GAT_CHUNK edges is cut into chunks of GODE_GAT_CHUNK consecutive
 * edges that separate threads reduce; its chunks are added in chunk order (deterministic).  n_chunks == 0: no hubs. */
#define GODE_GAT_CHUNK 64
typedef struct {
  int32_t n_heavy, n_chunks;
  const int32_t* nodes;       /* [n_heavy] */
  const int32_t* cptr;        /* [n_heavy + 1] chunk range of each hub */
  const int32_t* chunk_node;  /* [n_chunks] */
  const int32_t* chunk_e0;    /* [n_chunks] first edge of the chunk (position in the grouping) */
} gode_gat_heavy_t;

typedef struct {
  int64_t n_nodes;
  int64_t n_edges;
  const int32_t* tptr;   /* [n_nodes + 1] segments by target */
  const int32_t* t_src;  /* [n_edges] source node of each edge, by-target order */
  const int32_t* t_tgt;  /* [n_edges] target node of each edge, by-target order */
  const int32_t* sptr;   /* [n_nodes + 1] segments by source */
  const int32_t* s_tgt;  /* [n_edges] target node of each edge, by-source order */
  const int32_t* s_pos;  /* [n_edges] by-target position of each edge, by-source order */
  gode_gat_heavy_t t_heavy;  /* hubs of the by-target grouping */
  gode_gat_heavy_t s_heavy;  /* hubs of the by-source grouping */
} gode_gat_graph_t;

size_t gode_gat_fwd_workspace_bytes(const gode_gat_graph_t* g, int32_t heads, int32_t oh);
int gode_gat_fwd(const gode_gat_graph_t* g, int32_t heads, int32_t oh, const float* P, int64_t ldp, float eps,
                 float* out, int64_t ldo, float* den, unsigned long long* amax_key, int32_t* nan_flag,
                 void* ws, size_t ws_bytes, void* stream);
size_t gode_gat_bwd_workspace_bytes(const gode_gat_graph_t* g, int32_t heads, int32_t oh);
int gode_gat_bwd(const gode_gat_graph_t* g, int32_t heads, int32_t oh, const float* P, int64_t ldp,
                 const float* out, int64_t ldo, const float* den, const unsigned long long* amax_key,
                 const float* gout, int64_t ldg, float* dP, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * QC edge-conditioned messages: the per-edge mat-vec.
 * replaces: index_select(support, 0, Esrc) + torch.bmm(edge_data, .) -- QC/layers.py:143-144, QC/mpnn.py:27-28.
 *   msg[e, r] = sum_c edge_data[e, r, c] * s[esrc[e], c]        edge_data [E, f, f] contiguous, f <= 128
 * The reference's dense one-hot product torch.spmm(Etgt, edge_msg) (QC/layers.py:145) is the segmented sum
 * gode_spmm_csr_f32 over msg with the CSR of Etgt (rows = targets, columns = edge ids, values 1).
 * bwd (dm_e = gout[etgt[e]], or gout[e] when etgt is NULL):
 *                              ds_msg[e, c] = sum_r edge_data[e, r, c] * dm_e[r]   (scatter by esrc afterwards)
 *                              d_edge_data[e, r, c] = dm_e[r] * s[esrc[e], c]       (skipped when NULL)
 * ---------------------------------------------------------------------------------------------- */
int gode_edge_matvec(int64_t n_edges, int32_t f, const float* edge_data, const int32_t* esrc,
                     const float* s, int64_t lds, float* msg, void* stream);
int gode_edge_matvec_bwd(int64_t n_edges, int32_t f, const float* edge_data, const int32_t* esrc,
                         const int32_t* etgt, const float* s, int64_t lds, const float* gout, int64_t ldg,
                         float* ds_msg, float* d_edge_data, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attention readout over the nodes of each graph of a batch: the inner step of Set2Set.
 * replaces: QC/set2set.py:60-75 (e = <x_i, q[batch_i]>, per-graph softmax in a Python loop over the batch with boolean
 *           masks, scatter_add of a * x).  Nodes of graph g are rows gptr[g] .. gptr[g+1] of x (block-diagonal batches).
 *   fwd: a[i] = softmax_g(<x_i, q_g>)  (a [N] is saved for the backward),  r[g, :] = sum_i a[i] x[i, :]
 *   bwd: dx [N, h] and dq [B, h] from dr [B, h].   h <= 256.  A graph without nodes gives r = 0.
 * ---------------------------------------------------------------------------------------------- */
int gode_segment_attend_fwd(int32_t n_graphs, int32_t h, const int32_t* gptr, const float* x, int64_t ldx,
                            const float* q, int64_t ldq, float* a, float* r, int64_t ldr, void* stream);
int gode_segment_attend_bwd(int32_t n_graphs, int32_t h, const int32_t* gptr, const float* x, int64_t ldx,
                            const float* q, int64_t ldq, const float* a, const float* dr, int64_t lddr,
                            float* dx, int64_t lddx, float* dq, int64_t lddq, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Device-side collate of molecule graphs into one block-diagonal batch.
 * replaces: collate_g_concat_edge_data -- QC/datasets/utils.py:153-217 (the DataLoader's collate_fn, a Python loop per
 *           batch): node ids shifted by the atoms before the molecule, every undirected edge emitted as
 *           (src -> tgt) at m_acc + i and (tgt -> src) at M + m_acc + i with the same feature row, B[j] = position of
 *           node j's molecule in the batch.
 * The dataset is a ragged store on the device: node_ptr / edge_ptr [n_molecules + 1] (int64), X_all [sum atoms, n_d],
 * per undirected edge (in the reference's sorted((src, tgt)) order inside a molecule) local node ids e_src_local /
 * e_tgt_local (int32) and E_all [sum edges, e_d].  sel [n_sel] (int32) are the molecule ids of the batch in batch order;
 * node_off / edge_off [n_sel + 1] the exclusive scans of their atom / edge counts, N = node_off[n_sel], M = edge_off[n_sel].
 * Outputs: B [N] int64, X [N, n_d], E_d [2M, e_d], E_src [2M] int64, E_tgt [2M] int64 (the target node of every
 * directed edge: the row index of the reference's one-hot E_tgt [N, 2M]).  shift = 0 keeps the endpoints as numbered
 * inside the molecule (the reference, utils.py:196-203); shift = 1 adds the molecule's node offset in the batch.
 * ---------------------------------------------------------------------------------------------- */
int gode_qc_collate(int32_t n_sel, const int32_t* sel, const int64_t* node_ptr, const int64_t* edge_ptr,
                    const int64_t* node_off, const int64_t* edge_off, int32_t shift, int64_t N, int64_t M,
                    const float* X_all, int32_t n_d, const int32_t* e_src_local, const int32_t* e_tgt_local, const float* E_all, int32_t e_d,
                    int64_t* B, float* X, float* E_d, int64_t* E_src, int64_t* E_tgt, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Halo pack for the row-partitioned (multi-GPU) path: dst[i, :] = src[idx[i], :].
 * The reference is single-device; this packs the rows of the SpMM operand of
 * torch.spmm(adj, support) (GCN/layers.py:71) that peer ranks reference, ordered by destination rank,
 * for the NCCL all-to-all-v issued by the host code (graph-odenet_b200/parallel.py).
 * ---------------------------------------------------------------------------------------------- */
int gode_gather_rows(int64_t n_idx, const int32_t* idx, int32_t d, const float* src, int64_t lds,
                     float* dst, int64_t ldd, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Halo exchange over NVLink peer memory (csrc/peer.cu): one kernel reads the rows peers reference from the
 * local operand and stores them straight into the halo tails of the peers' operand buffers, then publishes
 * an epoch flag; the consumer waits for every peer's flag.  Replaces gode_gather_rows + NCCL all-to-all-v for
 * the same step of the row-partitioned torch.spmm(adj, support) (GCN/layers.py:71).
 *
 * Every rank owns one arena (gode_peer_alloc: cudaMalloc, first GODE_PEER_HEADER_BYTES zeroed and reserved:
 * uint32 flags[GODE_MAX_PEERS] at offset 0, CTA counter at GODE_PEER_COUNTER_OFFSET, status word at
 * GODE_PEER_STATUS_OFFSET).  Operand buffers are carved from the arena at the SAME offsets on every rank.
 * gode_peer_export / gode_peer_open move the 64-byte cudaIpc handle between processes (the host code ships the
 * bytes, e.g. with torch.distributed.all_gather_object).  base[rank] is the local arena, base[p] the mapping of
 * peer p's arena.  Epochs are the exchange counter (1, 2, ...; wrap-safe), identical on every rank.
 * gode_halo_push: send_idx[send_ptr[p] .. send_ptr[p+1]) are the local rows peer p needs (host array send_ptr of
 * world+1 entries); they land at rows dst_row[p] + k of the buffer at byte offset buf_offset of p's arena (leading
 * dimension ldd floats).  No host synchronisation.  gode_peer_wait spins on the device until every peer has
 * published `epoch`, at most timeout_ns; on time-out the status word becomes GODE_PEER_TIMEOUT (read it with
 * gode_peer_status, which synchronises the stream).
 * ---------------------------------------------------------------------------------------------- */
#define GODE_PEER_HANDLE_BYTES 64
#define GODE_PEER_HEADER_BYTES 4096
#define GODE_PEER_COUNTER_OFFSET 128
#define GODE_PEER_STATUS_OFFSET 192
#define GODE_PEER_TIMEOUT 1

typedef struct {
  int32_t world, rank;
  void* base[GODE_MAX_PEERS];
} gode_peer_group_t;

int gode_peer_alloc(size_t bytes, void** out);
int gode_peer_free(void* p);
int gode_peer_export(const void* p, void* handle64);
int gode_peer_open(const void* handle64, void** out);
int gode_peer_close(void* p);
int gode_halo_push(const gode_peer_group_t* g, uint32_t epoch, const int32_t* send_idx, const int64_t* send_ptr,
                   const int64_t* dst_row, int64_t buf_offset, int32_t d, const float* src, int64_t lds,
                   int64_t ldd, int32_t max_ctas, void* stream);
/* One part of a pipelined exchange: the entries send_idx[seg_begin[p] .. seg_end[p]) of peer p go to rows dst_row[p] + k;
 * only the call with signal != 0 (the last part, issued on the same stream) publishes the epoch. */
int gode_halo_push_part(const gode_peer_group_t* g, uint32_t epoch, const int32_t* send_idx,
                        const int64_t* seg_begin, const int64_t* seg_end, const int64_t* dst_row,
                        int64_t buf_offset, int32_t d, const float* src, int64_t lds, int64_t ldd,
                        int32_t max_ctas, int32_t signal, void* stream);
int gode_peer_wait(const gode_peer_group_t* g, uint32_t epoch, uint64_t timeout_ns, void* stream);
int gode_peer_status(const gode_peer_group_t* g, int32_t* status_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GODE_H_ */
