"""Tensor-level wrappers over the C ABI (libgode.so) + autograd Functions used by the layer modules.

PyTorch is plumbing here: it owns device memory (caching allocator) and the current stream; every
arithmetic step of the hot path is a libgode kernel.  Nothing in this file falls back to ATen math on
the hot path -- tensors that are not CUDA fp32 raise.
"""
from __future__ import annotations

import ctypes as C
import weakref

import torch

from . import _lib
from ._lib import lib, check

_NULL = None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _req(t, name, dtype=torch.float32):
    if not (torch.is_tensor(t) and t.is_cuda and t.dtype == dtype):
        raise TypeError("%s must be a CUDA %s tensor (got %s); graph-odenet_b200 has no CPU path" % (
            name, dtype, (t.device, t.dtype) if torch.is_tensor(t) else type(t)))
    return t


def _rowmajor(t, name):
    _req(t, name)
    if t.dim() != 2 or t.stride(1) != 1:
        t = t.contiguous()
    return t


_ws_cache = {}


def workspace(nbytes, device, tag="ws"):
    """Grow-only per-(device, tag) scratch buffer from the caching allocator."""
    key = (device.index, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


# --------------------------------------------------------------------------------------------------
# graph plan
# --------------------------------------------------------------------------------------------------


class GraphPlan:
    """Canonical CSR and CSR^T (int32) of an adjacency, built once on the device.

    replaces the per-call COO handling inside ``torch.spmm`` (GCN/layers.py:33,71).  ``from_coo`` accepts
    what the reference loader produces (GCN/utils.py:222-229): int64 indices [2, nnz], fp32 values,
    uncoalesced, unsorted columns.
    """

    def __init__(self, n_rows, n_cols, rowptr, colidx, vals, build_transpose=True):
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.rowptr, self.colidx, self.vals = rowptr, colidx, vals
        self.nnz = int(colidx.numel())
        self.device = rowptr.device
        self.heavy, self.chunk_ptr, self.n_heavy, self.n_chunks = self._heavy(rowptr, self.n_rows)
        self.row_vals = self._row_values(rowptr, vals, self.n_rows)
        self.row_vals_t = None
        self.rowptr_t = self.colidx_t = self.vals_t = self.perm_t = None
        self.heavy_t, self.chunk_ptr_t, self.n_heavy_t, self.n_chunks_t = None, None, 0, 0
        if build_transpose:
            self._build_transpose()
        self._csr = self._csr_t = None

    @staticmethod
    def _heavy(rowptr, n_rows):
        """Heavy (hub) rows and the chunk table the SpMM splits them by (gode_csr_heavy_rows)."""
        if n_rows == 0:
            return None, None, 0, 0
        rows = torch.empty(n_rows, dtype=torch.int32, device=rowptr.device)
        cptr = torch.empty(n_rows + 1, dtype=torch.int32, device=rowptr.device)
        cnt = torch.zeros(2, dtype=torch.int32, device=rowptr.device)
        check(lib.gode_csr_heavy_rows(n_rows, _p(rowptr), _p(rows), _p(cptr), _p(cnt), _stream()), "gode_csr_heavy_rows")
        n, nc = [int(v) for v in cnt.tolist()]
        if n == 0:
            return None, None, 0, 0
        return rows[:n].clone(), cptr[:n + 1].clone(), n, nc

    @staticmethod
    def _row_values(rowptr, vals, n_rows):
        """Per-row common value if the matrix is row-constant (D^-1 (A+I) is), else None (gode_csr_row_values)."""
        if n_rows == 0 or vals.numel() == 0:
            return None
        out = torch.empty(n_rows, dtype=torch.float32, device=rowptr.device)
        flag = torch.zeros(1, dtype=torch.int32, device=rowptr.device)
        check(lib.gode_csr_row_values(n_rows, _p(rowptr), _p(vals), _p(out), _p(flag), _stream()), "gode_csr_row_values")
        return out if int(flag.item()) == 1 else None

    def csr(self, transpose=False):
        """The ``gode_csr_t`` view of A (or A^T) passed to the kernels."""
        if transpose:
            if self._csr_t is None:
                if self.rowptr_t is None:
                    raise RuntimeError("this plan was built without its transpose")
                self._csr_t = _make_csr(self.n_cols, self.n_rows, self.rowptr_t, self.colidx_t, self.vals_t,
                                        self.heavy_t, self.chunk_ptr_t, self.n_heavy_t, self.n_chunks_t, self.row_vals_t)
            return self._csr_t
        if self._csr is None:
            self._csr = _make_csr(self.n_rows, self.n_cols, self.rowptr, self.colidx, self.vals, self.heavy,
                                  self.chunk_ptr, self.n_heavy, self.n_chunks, self.row_vals)
        return self._csr

    def row_stochastic_scale(self):
        """``row_vals`` when the matrix is row-constant AND every row sums to one (the reference's D^-1 (A + I),
        GCN/utils.py:205-212: row value = fp32(1 / row count)), else None.  Decides ``unit_transpose``."""
        if not hasattr(self, "_rs_scale"):
            self._rs_scale = None
            if self.row_vals is not None and self.n_rows > 0:
                cnt = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.float32)
                if bool(((self.row_vals * cnt - 1.0).abs() <= 1e-6).all().item()):
                    self._rs_scale = self.row_vals
        return self._rs_scale

    def unit_pattern_csr(self):
        """``gode_csr_t`` of this matrix's 0/1 PATTERN (same index arrays, every value one, row-constant)."""
        if getattr(self, "_unit_csr", None) is None:
            self._unit_vals = torch.ones(max(self.nnz, 1), dtype=torch.float32, device=self.device)
            self._unit_row_vals = torch.ones(max(self.n_rows, 1), dtype=torch.float32, device=self.device)
            self._unit_csr = _make_csr(self.n_rows, self.n_cols, self.rowptr, self.colidx, self._unit_vals, self.heavy,
                                       self.chunk_ptr, self.n_heavy, self.n_chunks, self._unit_row_vals)
        return self._unit_csr

    def unit_transpose(self):
        """(row scale, pattern CSR of A^T) when the adjoint may gather ``A_hat^T gP`` without a values stream: for a
        row-constant, row-stochastic A_hat column i of A_hat^T is the constant rv_i, so the producer of gP stores
        ``rv_i * gP_i`` and the gather adds plain rows; the bias gradient ``sum_i gP_i`` equals the column sums of the
        gathered result because every row of A_hat sums to one (gode_gcn_odefunc_t.gp_row_scale).  Else None."""
        if self.rowptr_t is None or self.n_rows != self.n_cols:
            return None
        scale = self.row_stochastic_scale()
        if scale is None:
            return None
        if getattr(self, "_unit_t_view", None) is None:
            self._unit_t_view = self.transposed()      # shares the index arrays; owns the unit values
        return scale, self._unit_t_view.unit_pattern_csr()

    def _build_transpose(self):
        dev = self.device
        nnz = self.nnz
        self.rowptr_t = torch.empty(self.n_cols + 1, dtype=torch.int32, device=dev)
        self.colidx_t = torch.empty(nnz, dtype=torch.int32, device=dev)
        self.vals_t = torch.empty(nnz, dtype=torch.float32, device=dev)
        self.perm_t = torch.empty(nnz, dtype=torch.int32, device=dev)
        nb = lib.gode_csr_transpose_workspace_bytes(nnz, self.n_rows, self.n_cols)
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        check(lib.gode_csr_transpose(self.n_rows, self.n_cols, nnz, _p(self.rowptr), _p(self.colidx), _p(self.vals),
                                     _p(self.rowptr_t), _p(self.colidx_t), _p(self.vals_t), _p(self.perm_t),
                                     _p(ws), nb, _stream()), "gode_csr_transpose")
        self.heavy_t, self.chunk_ptr_t, self.n_heavy_t, self.n_chunks_t = self._heavy(self.rowptr_t, self.n_cols)
        self.row_vals_t = self._row_values(self.rowptr_t, self.vals_t, self.n_cols)
        del ws

    @classmethod
    def from_coo(cls, row, col, val, n_rows, n_cols, build_transpose=True):
        dev = val.device
        if not val.is_cuda:
            raise TypeError("GraphPlan.from_coo needs CUDA tensors")
        row = row.to(torch.int64).contiguous()
        col = col.to(torch.int64).contiguous()
        val = val.to(torch.float32).contiguous()
        nnz = int(val.numel())
        rowptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
        colidx = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        vals = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
        nnz_out = torch.zeros(1, dtype=torch.int64, device=dev)
        nb = lib.gode_csr_from_coo_workspace_bytes(nnz, n_rows)
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        check(lib.gode_csr_from_coo(n_rows, n_cols, nnz, _p(row), _p(col), _p(val), _p(rowptr), _p(colidx), _p(vals),
                                    _p(nnz_out), _p(ws), nb, _stream()), "gode_csr_from_coo")
        n_u = int(nnz_out.item())
        if n_u < 0:
            raise IndexError("adjacency index out of range for shape (%d, %d)" % (n_rows, n_cols))
        del ws
        return cls(n_rows, n_cols, rowptr, colidx[:n_u].clone() if n_u < nnz else colidx[:n_u],
                   vals[:n_u].clone() if n_u < nnz else vals[:n_u], build_transpose)

    @classmethod
    def from_sparse(cls, adj, build_transpose=True):
        """``adj``: torch sparse COO (as the reference passes) or a dense matrix (GCN-dense-paper passes one)."""
        if adj.layout == torch.sparse_coo:
            idx, val = adj._indices(), adj._values()
            return cls.from_coo(idx[0], idx[1], val, adj.shape[0], adj.shape[1], build_transpose)
        if adj.layout == torch.strided and adj.dim() == 2:
            idx = adj.nonzero(as_tuple=False).t()
            return cls.from_coo(idx[0], idx[1], adj[idx[0], idx[1]], adj.shape[0], adj.shape[1], build_transpose)
        raise TypeError("unsupported adjacency layout %s" % adj.layout)

    def transposed(self):
        """Plan of A^T sharing storage (used by the backward of the discrete layers)."""
        t = object.__new__(GraphPlan)
        t.n_rows, t.n_cols, t.nnz, t.device = self.n_cols, self.n_rows, self.nnz, self.device
        t.rowptr, t.colidx, t.vals = self.rowptr_t, self.colidx_t, self.vals_t
        t.heavy, t.chunk_ptr, t.n_heavy, t.n_chunks = self.heavy_t, self.chunk_ptr_t, self.n_heavy_t, self.n_chunks_t
        t.rowptr_t, t.colidx_t, t.vals_t, t.perm_t = self.rowptr, self.colidx, self.vals, None
        t.heavy_t, t.chunk_ptr_t, t.n_heavy_t, t.n_chunks_t = self.heavy, self.chunk_ptr, self.n_heavy, self.n_chunks
        t.row_vals, t.row_vals_t = self.row_vals_t, self.row_vals
        t._csr = t._csr_t = None
        return t


def tile_schedule(n_rows, heavy, chunk_ptr, n_heavy, n_chunks):
    """Work order of the d = 128 tile gather (gode_csr_t.tile_sched): the 32-row tiles in id order, every group of 8 hub
    chunks right behind the tile that contains the hub of its first chunk.  int32 on the device, or None without hubs."""
    if not n_heavy or n_rows == 0:
        return None
    dev = heavy.device
    n_tiles = (n_rows + 31) // 32
    groups = (n_chunks + 7) // 8
    first = torch.arange(groups, device=dev, dtype=torch.int64) * 8
    hub = torch.searchsorted(chunk_ptr[:n_heavy + 1].to(torch.int64), first, right=True) - 1
    row = heavy[hub.clamp_(0, n_heavy - 1)].to(torch.int64)
    keys = torch.cat([torch.arange(n_tiles, device=dev, dtype=torch.int64) * 2, (row // 32) * 2 + 1])
    items = torch.cat([torch.arange(n_tiles, device=dev, dtype=torch.int64),
                       -(torch.arange(groups, device=dev, dtype=torch.int64) + 1)])
    order = torch.sort(keys, stable=True).indices
    return items[order].to(torch.int32).contiguous()


def _make_csr(n_rows, n_cols, rowptr, colidx, vals, heavy, chunk_ptr, n_heavy, n_chunks, row_vals=None):
    c = _lib.Csr()
    sched = tile_schedule(n_rows, heavy, chunk_ptr, n_heavy, n_chunks)
    if sched is not None:
        c._keep_sched = sched        # the struct instance is cached on its plan: the table lives as long as the plan
        c.tile_sched, c.n_tile_sched = sched.data_ptr(), int(sched.numel())
    c.n_rows, c.n_cols = n_rows, n_cols
    c.rowptr, c.colidx, c.vals = rowptr.data_ptr(), colidx.data_ptr(), vals.data_ptr()
    c.row_vals = row_vals.data_ptr() if row_vals is not None else None
    c.heavy_rows = heavy.data_ptr() if n_heavy else None
    c.heavy_chunk_ptr = chunk_ptr.data_ptr() if n_heavy else None
    c.n_heavy, c.n_chunks = n_heavy, n_chunks
    return c


_plan_cache = {}


def plan_for(adj):
    """Plan cache keyed on the adjacency tensor's storage (the reference passes the same ``adj`` every epoch)."""
    if isinstance(adj, GraphPlan):
        return adj
    if adj.layout == torch.sparse_coo:
        key = (adj._indices().data_ptr(), adj._values().data_ptr(), tuple(adj.shape), adj._nnz(), adj.device.index)
    else:
        key = (adj.data_ptr(), 0, tuple(adj.shape), adj.numel(), adj.device.index)
    hit = _plan_cache.get(key)
    if hit is not None and hit[0]() is adj:
        return hit[1]
    if not adj.is_cuda:
        raise TypeError("adjacency must live on a CUDA device (call adj.cuda() as GCN/train_res.py:57 does)")
    plan = GraphPlan.from_sparse(adj)
    try:
        _plan_cache[key] = (weakref.ref(adj), plan)
    except TypeError:
        pass
    if len(_plan_cache) > 64:
        for k in [k for k, v in _plan_cache.items() if v[0]() is None]:
            _plan_cache.pop(k, None)
    return plan


# --------------------------------------------------------------------------------------------------
# raw kernels
# --------------------------------------------------------------------------------------------------


def spmm(plan, x, bias=None, relu=False, residual=None, out=None, transpose=False):
    """out = relu?(A x + bias) + residual  via gode_spmm_csr_f32."""
    x = _rowmajor(x, "x")
    csr = plan.csr(transpose)
    if x.shape[0] != csr.n_cols:
        raise ValueError("spmm: operand has %d rows, adjacency expects %d" % (x.shape[0], csr.n_cols))
    d = x.shape[1]
    if out is None:
        out = torch.empty(csr.n_rows, d, dtype=torch.float32, device=x.device)
    ep = _lib.SpmmEpilogue()
    ep.bias = _p(_req(bias, "bias").contiguous()) if bias is not None else None
    ep.relu = 1 if relu else 0
    if residual is not None:
        residual = _rowmajor(residual, "residual")
        if residual.stride(0) != out.stride(0):
            residual = residual.contiguous()
        ep.residual = _p(residual)
    nb = lib.gode_spmm_workspace_bytes(C.byref(csr), d)
    ws = workspace(nb, x.device, "spmm") if nb else None
    check(lib.gode_spmm_csr_f32(C.byref(csr), _p(x), x.stride(0), d, _p(out), out.stride(0), C.byref(ep), _p(ws), nb,
                                _stream()), "gode_spmm_csr_f32")
    return out


def gemm(a, b, trans_a=False, trans_b=False, out=None, alpha=1.0, beta=0.0, splits=1):
    """out = alpha * op(a) @ op(b) + beta * out  via gode_gemm_f32."""
    a = _rowmajor(a, "a")
    b = _rowmajor(b, "b")
    M, K = (a.shape[1], a.shape[0]) if trans_a else a.shape
    K2, N = (b.shape[1], b.shape[0]) if trans_b else b.shape
    if K != K2:
        raise ValueError("gemm: inner dimensions differ (%d vs %d)" % (K, K2))
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
        beta = 0.0
    ws, nb = None, 0
    if splits > 1:
        nb = 4 * splits * M * N
        ws = workspace(nb, a.device, "gemm")
    check(lib.gode_gemm_f32(int(trans_a), int(trans_b), M, N, K, alpha, _p(a), a.stride(0), _p(b), b.stride(0), beta,
                            _p(out), out.stride(0), _lib.PREC_FP32, splits, _p(ws), nb, _stream()), "gode_gemm_f32")
    return out


def _splits_for(k, m, n):
    if k < 8192 or m * n > 512 * 512:
        return 1
    return int(min(k // 4096, 256))


def colsum(x):
    x = _rowmajor(x, "x")
    out = torch.empty(x.shape[1], dtype=torch.float32, device=x.device)
    nb = lib.gode_colreduce_workspace_bytes(x.shape[1])
    ws = workspace(nb, x.device, "colreduce")
    check(lib.gode_colsum_f32(x.shape[0], x.shape[1], _p(x), x.stride(0), _p(out), _p(ws), nb, _stream()), "gode_colsum_f32")
    return out


def groupnorm_fwd(x, groups, gamma, beta, eps):
    x = _rowmajor(x, "x")
    y = torch.empty_like(x, memory_format=torch.contiguous_format)
    check(lib.gode_groupnorm_fwd(x.shape[0], x.shape[1], groups, eps, _p(x), x.stride(0), _p(gamma.contiguous()),
                                 _p(beta.contiguous()), _p(y), y.stride(0), _stream()), "gode_groupnorm_fwd")
    return y


def groupnorm_bwd(x, groups, gamma, dy, eps):
    x = _rowmajor(x, "x")
    dy = _rowmajor(dy, "dy")
    d = x.shape[1]
    dx = torch.empty_like(x, memory_format=torch.contiguous_format)
    dg = torch.empty(d, dtype=torch.float32, device=x.device)
    db = torch.empty(d, dtype=torch.float32, device=x.device)
    nb = lib.gode_colreduce_workspace_bytes(2 * d)
    ws = workspace(nb, x.device, "colreduce")
    check(lib.gode_groupnorm_bwd(x.shape[0], d, groups, eps, _p(x), x.stride(0), _p(gamma.contiguous()), _p(dy),
                                 dy.stride(0), _p(dx), dx.stride(0), _p(dg), _p(db), _p(ws), nb, _stream()),
          "gode_groupnorm_bwd")
    return dx, dg, db


def rk_combine(y0, ks, coefs, out=None):
    """out = y0 + sum_j coefs[j] * ks[j]   (y0 may be None)."""
    ref = ks[0] if y0 is None else y0
    if out is None:
        out = torch.empty_like(ref)
    karr = (C.c_void_p * _lib.MAX_STAGES)(*[k.data_ptr() for k in ks])
    carr = (C.c_float * _lib.MAX_STAGES)(*[float(c) for c in coefs])
    check(lib.gode_rk_combine(ref.numel(), _p(y0), karr, carr, len(ks), _p(out), _stream()), "gode_rk_combine")
    return out


def rk_error_sumsq(y0, y1, ks, coefs, rtol, atol):
    """Device scalar: sum_i (sum_j c_j k_j[i] / (atol + rtol*max(|y0_i|,|y1_i|)))^2."""
    out = torch.empty(1, dtype=torch.float32, device=y0.device)
    nb = lib.gode_colreduce_workspace_bytes(1)
    ws = workspace(nb, y0.device, "colreduce")
    karr = (C.c_void_p * _lib.MAX_STAGES)(*[k.data_ptr() for k in ks])
    carr = (C.c_float * _lib.MAX_STAGES)(*[float(c) for c in coefs])
    check(lib.gode_rk_error_sumsq(y0.numel(), _p(y0), _p(y1), karr, carr, len(ks), rtol, atol, _p(out), _p(ws), nb,
                                  _stream()), "gode_rk_error_sumsq")
    return out


# --------------------------------------------------------------------------------------------------
# autograd Functions for the discrete layers
# --------------------------------------------------------------------------------------------------


class GraphConvFn(torch.autograd.Function):
    """``spmm(adj, mm(x, W)) + b`` (GCN/layers.py:31-37) and its backward, all on libgode kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias, plan, relu):
        x = _rowmajor(x, "input")
        w = _rowmajor(weight, "weight")
        support = gemm(x, w)
        out = spmm(plan, support, bias=bias, relu=relu)
        ctx.plan, ctx.relu, ctx.has_bias = plan, relu, bias is not None
        ctx.save_for_backward(x, w, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w, out = ctx.saved_tensors
        g = _rowmajor(g, "grad")
        if ctx.relu:
            g = relu_mask(g, out)
        gb = colsum(g) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        gs = spmm(ctx.plan, g, transpose=True)                       # A^T g
        gx = gemm(gs, w, trans_b=True) if ctx.needs_input_grad[0] else None
        gw = None
        if ctx.needs_input_grad[1]:
            gw = gemm(x, gs, trans_a=True, splits=_splits_for(x.shape[0], x.shape[1], gs.shape[1]))
        return gx, gw, gb, None, None


class GroupNormFn(torch.autograd.Function):
    """nn.GroupNorm on [N, d] (GCN/models.py:88,129,143,165)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps):
        ctx.groups, ctx.eps = groups, eps
        ctx.save_for_backward(x, gamma)
        return groupnorm_fwd(x, groups, gamma, beta, eps)

    @staticmethod
    def backward(ctx, dy):
        x, gamma = ctx.saved_tensors
        dx, dg, db = groupnorm_bwd(x, ctx.groups, gamma, dy, ctx.eps)
        return dx, dg, db, None, None


class SparseInputGraphConvFn(torch.autograd.Function):
    """``spmm(adj, mm(X, W)) + b`` for a SPARSE feature matrix X: the reference densifies its
    bag-of-words features (GCN/utils.py:192, ~1 % non-zero) and runs a dense [N, F] x [F, h] product
    (GCN/layers.py:32); here CSR(X) W is the same gather kernel as the adjacency product, with W as the gathered
    operand, and the weight gradient is CSR(X)^T (A_hat^T g).  X is a constant: it gets no gradient."""

    @staticmethod
    def forward(ctx, weight, bias, plan_x, plan, relu):
        w = _rowmajor(weight, "weight")
        support = spmm(plan_x, w)
        out = spmm(plan, support, bias=bias, relu=relu)
        ctx.plan_x, ctx.plan, ctx.relu, ctx.has_bias = plan_x, plan, relu, bias is not None
        ctx.save_for_backward(out if relu else None)
        return out

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        g = _rowmajor(g, "grad")
        if ctx.relu:
            g = relu_mask(g, out)
        gb = colsum(g) if (ctx.has_bias and ctx.needs_input_grad[1]) else None
        gw = None
        if ctx.needs_input_grad[0]:
            gs = spmm(ctx.plan, g, transpose=True)                   # A^T g
            gw = spmm(ctx.plan_x, gs, transpose=True)                # X^T (A^T g)
        return gw, gb, None, None, None


def _is_sparse(x):
    return isinstance(x, torch.Tensor) and x.layout == torch.sparse_coo


def graph_conv(x, adj, weight, bias, relu=False):
    """``x``: dense [N, F] features as the reference passes them, or -- extension -- a sparse COO tensor / a GraphPlan
    of the feature matrix."""
    if _is_sparse(x) or isinstance(x, GraphPlan):
        if _is_sparse(x) and x.requires_grad:
            raise ValueError("sparse input features are constants: requires_grad is not supported")
        return SparseInputGraphConvFn.apply(weight, bias, plan_for(x), plan_for(adj), relu)
    return GraphConvFn.apply(x, weight, bias, plan_for(adj), relu)


def group_norm(x, groups, gamma, beta, eps=1e-5):
    return GroupNormFn.apply(x, gamma, beta, groups, eps)


# --------------------------------------------------------------------------------------------------
# Linear with fused bias / ReLU (GAT projections, QC MyLinear / MLP)
# --------------------------------------------------------------------------------------------------


def _tc_gemm_enabled(M, N, K):
    """Products large enough to fill 128 x 128 tiles go to the general tcgen05 GEMM (gode_gemm_tc_f32: 3xTF32 with
    hierarchical accumulation, as accurate as an fp32 SGEMM -- tests/test_gpu_gemm_tc.py); GODE_GEMM_TC=0 keeps everything
    on the SIMT kernel."""
    import os
    return os.environ.get("GODE_GEMM_TC", "1") != "0" and M >= 512 and N >= 64 and K >= 32


def _pad4(t):
    """Row-major copy of a 2-D tensor whose leading dimension is a multiple of 4 floats (128-bit accesses); a view of the
    padded buffer with the original shape."""
    ld = (t.shape[1] + 3) // 4 * 4
    if t.stride(1) == 1 and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0:
        return t
    buf = torch.zeros(t.shape[0], ld, dtype=torch.float32, device=t.device)
    buf[:, :t.shape[1]].copy_(t)
    return buf[:, :t.shape[1]]


def gemm_tc(a, bt, bias=None, relu=False, out=None):
    """``act(a @ bt.T + bias)`` with both operands K-major -- gode_gemm_tc_f32."""
    a, bt = _pad4(_rowmajor(a, "a")), _pad4(_rowmajor(bt, "bt"))
    M, K = a.shape
    N = bt.shape[0]
    if bt.shape[1] != K:
        raise ValueError("gemm_tc: inner dimensions differ")
    if out is None:
        # contiguous [M, N]: callers pass the result's width as its leading dimension (N % 4 != 0 -> the kernel's epilogue
        # falls back to 32-bit stores, which costs less than a padded buffer plus a compacting copy)
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    b = _req(bias, "bias").contiguous() if bias is not None else None
    check(lib.gode_gemm_tc_f32(M, N, K, _p(a), a.stride(0), _p(bt), bt.stride(0), _p(b), int(relu), _p(out), out.stride(0),
                               _lib.PREC_FP32, _stream()), "gode_gemm_tc_f32")
    return out


def gemm_tc_reduce_rows(a, b):
    """``a.T @ b`` for two [R, *] row-major operands with R large (weight gradients x^T g): both are transposed once
    (two copies, ~1 % of the product's traffic), the reduction over R runs as a split-K tcgen05 product over all SMs, and
    the slabs' partial products are added in slab order."""
    at, bt = _pad4(a.t().contiguous()), _pad4(b.t().contiguous())       # [M, R], [N, R]: K-major over the rows
    M, K = at.shape
    N = bt.shape[0]
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    sm = torch.cuda.get_device_properties(a.device).multi_processor_count
    splits = max(1, min((K + 2047) // 2048, (2 * sm + tiles - 1) // tiles))
    part = torch.empty(splits, M, N, dtype=torch.float32, device=a.device)
    check(lib.gode_gemm_tc_splitk_f32(M, N, K, _p(at), at.stride(0), _p(bt), bt.stride(0), _p(part), N, splits,
                                      _lib.PREC_FP32, _stream()), "gode_gemm_tc_splitk_f32")
    return part[0] if splits == 1 else part.sum(0)


def linear(x, weight, bias=None, relu=False, weight_is_out_in=False, out=None):
    """``act(x @ W + b)`` via gode_linear_f32.  ``weight`` is [in, out] (QC MyLinear) or, with
    ``weight_is_out_in``, [out, in] (nn.Linear)."""
    x = _rowmajor(x, "x")
    w = _rowmajor(weight, "weight")
    K = x.shape[1]
    N = w.shape[0] if weight_is_out_in else w.shape[1]
    if (w.shape[1] if weight_is_out_in else w.shape[0]) != K:
        raise ValueError("linear: inner dimensions differ")
    if out is None and _tc_gemm_enabled(x.shape[0], N, K):
        return gemm_tc(x, w if weight_is_out_in else w.t(), bias, relu)      # [in, out] weights are transposed once
    if out is None:
        out = torch.empty(x.shape[0], N, dtype=torch.float32, device=x.device)
    b = _req(bias, "bias").contiguous() if bias is not None else None
    check(lib.gode_linear_f32(int(weight_is_out_in), x.shape[0], N, K, _p(x), x.stride(0), _p(w), w.stride(0), _p(b),
                              int(relu), _p(out), out.stride(0), _stream()), "gode_linear_f32")
    return out


class LinearFn(torch.autograd.Function):
    """x @ W + b (+ ReLU) with W stored [in, out] -- QC/layers.py:26-30 (MyLinear) and the MLP's hidden ReLU (:55)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        x = _rowmajor(x, "x")
        w = _rowmajor(weight, "weight")
        out = linear(x, w, bias, relu)
        ctx.relu, ctx.has_bias = relu, bias is not None
        ctx.save_for_backward(x, w, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w, out = ctx.saved_tensors
        g = _rowmajor(g, "grad")
        if ctx.relu:
            g = relu_mask(g, out)
        M, K, N = x.shape[0], x.shape[1], g.shape[1]
        if _tc_gemm_enabled(M, min(K, N), min(K, N)):
            # dX = g W^T: g is K-major over N, W [K, N] is the "Bt" of that product as stored; dW = x^T g reduces over the
            # rows, so both operands are transposed first (two library copies, ~1 % of the product's time)
            gx = gemm_tc(g, w) if ctx.needs_input_grad[0] else None
            gw = gemm_tc_reduce_rows(x, g) if ctx.needs_input_grad[1] else None
        else:
            gx = gemm(g, w, trans_b=True) if ctx.needs_input_grad[0] else None
            gw = gemm(x, g, trans_a=True, splits=_splits_for(M, K, N)) if ctx.needs_input_grad[1] else None
        gb = colsum(g) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb, None


def relu_mask(g, out):
    """g * (out > 0) via gode_rk_combine-free path: one libgode elementwise kernel (gode_relu_bwd)."""
    g = g.contiguous()
    res = torch.empty_like(g)
    check(lib.gode_relu_bwd(g.numel(), _p(g), _p(out.contiguous()), _p(res), _stream()), "gode_relu_bwd")
    return res


# --------------------------------------------------------------------------------------------------
# GAT: edge groupings + attention aggregation (gode_gat_fwd / gode_gat_bwd)
# --------------------------------------------------------------------------------------------------


class GatGraph:
    """Edges grouped by target and by source (built once per edge list; replaces the N x E incidence ``Mtgt``
    of GAT/utils.py:194-196 -- the grouping by target *is* its CSR)."""

    def __init__(self, src, tgt, n_nodes):
        if not (src.is_cuda and tgt.is_cuda):
            raise TypeError("GAT edge lists must live on a CUDA device; graph-odenet_b200 has no CPU path")
        dev = src.device
        E = int(src.numel())
        self.n_nodes, self.n_edges, self.device = int(n_nodes), E, dev
        src64, tgt64 = src.to(torch.int64).contiguous(), tgt.to(torch.int64).contiguous()
        if E and (int(src64.min()) < 0 or int(tgt64.min()) < 0 or int(src64.max()) >= n_nodes or int(tgt64.max()) >= n_nodes):
            raise IndexError("edge endpoint out of range for %d nodes" % n_nodes)
        eid = torch.arange(E, device=dev, dtype=torch.int64)
        ones = torch.ones(E, dtype=torch.float32, device=dev)
        # CSR of Mtgt (rows = targets, columns = edge ids): gode_csr_from_coo keeps edge order inside a segment
        pt = GraphPlan.from_coo(tgt64, eid, ones, n_nodes, max(E, 1), build_transpose=False)
        ps = GraphPlan.from_coo(src64, eid, ones, n_nodes, max(E, 1), build_transpose=False)
        t_eid = pt.colidx[:E].to(torch.int64)
        s_eid = ps.colidx[:E].to(torch.int64)
        self.tptr, self.sptr = pt.rowptr, ps.rowptr
        self.t_src = src64[t_eid].to(torch.int32)
        self.t_tgt = tgt64[t_eid].to(torch.int32)
        self.s_tgt = tgt64[s_eid].to(torch.int32)
        pos = torch.empty(E, dtype=torch.int64, device=dev)
        pos[t_eid] = eid
        self.s_pos = pos[s_eid].to(torch.int32)
        g = _lib.GatGraph()
        g.n_nodes, g.n_edges = self.n_nodes, E
        g.tptr, g.t_src, g.t_tgt = self.tptr.data_ptr(), self.t_src.data_ptr(), self.t_tgt.data_ptr()
        g.sptr, g.s_tgt, g.s_pos = self.sptr.data_ptr(), self.s_tgt.data_ptr(), self.s_pos.data_ptr()
        self._hub_tables = [self._hubs(self.tptr, g.t_heavy), self._hubs(self.sptr, g.s_heavy)]   # keep the tensors alive
        self.c = g
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=dev)


    def _hubs(self, ptr, h):
        """Hub table of one grouping (gode_gat_heavy_t): nodes with more than GAT_CHUNK edges, cut into chunks."""
        lim = _lib.GAT_CHUNK
        deg = (ptr[1:] - ptr[:-1]).to(torch.int64)
        nodes = torch.nonzero(deg > lim).squeeze(1)
        if nodes.numel() == 0:
            h.n_heavy = h.n_chunks = 0
            return ()
        nch = (deg[nodes] + lim - 1) // lim
        cptr = torch.zeros(nodes.numel() + 1, dtype=torch.int64, device=ptr.device)
        cptr[1:] = torch.cumsum(nch, 0)
        chunk_node = torch.repeat_interleave(nodes, nch)
        local = torch.arange(int(cptr[-1]), device=ptr.device, dtype=torch.int64) - torch.repeat_interleave(cptr[:-1], nch)
        chunk_e0 = ptr[chunk_node].to(torch.int64) + local * lim
        t = tuple(x.to(torch.int32).contiguous() for x in (nodes, cptr, chunk_node, chunk_e0))
        h.n_heavy, h.n_chunks = int(nodes.numel()), int(chunk_node.numel())
        h.nodes, h.cptr, h.chunk_node, h.chunk_e0 = (x.data_ptr() for x in t)
        return t


_gat_cache = {}


def gat_graph_for(src, tgt, n_nodes):
    if isinstance(src, GatGraph):
        return src
    key = (src.data_ptr(), tgt.data_ptr(), int(src.numel()), int(n_nodes), src.device.index)
    hit = _gat_cache.get(key)
    if hit is not None and hit[0]() is src:
        return hit[1]
    g = GatGraph(src, tgt, n_nodes)
    _gat_cache[key] = (weakref.ref(src), g)
    if len(_gat_cache) > 64:
        for k in [k for k, v in _gat_cache.items() if v[0]() is None]:
            _gat_cache.pop(k, None)
    return g


class GatConvFn(torch.autograd.Function):
    """GAT/layers.py:40-58 on libgode: node projection GEMM -> gode_gat_fwd; backward gode_gat_bwd -> two GEMMs.

    ``f_weight`` [H*oh, 2i], ``f_bias`` [H*oh], ``w_weight`` [H, 2i], ``w_bias`` [H] (H = 1: the reference layer).
    """

    @staticmethod
    def forward(ctx, x, f_weight, f_bias, w_weight, w_bias, graph, heads, eps):
        x = _rowmajor(x, "x")
        i = x.shape[1]
        C_, H = f_weight.shape[0], heads
        oh = C_ // H
        pad = (-(2 * C_ + 2 * H)) % 4      # rows of P padded to a multiple of 16 bytes (zero weights): 128-bit kernels
        ldp = 2 * C_ + 2 * H + pad
        # W_cat [i, ldp] = [Wf_src^T | Wf_tgt^T | ww_src^T | ww_tgt^T | 0]; biases ride on the target halves
        zp = torch.zeros(i, pad, device=x.device)
        wcat = torch.cat([f_weight[:, :i].t(), f_weight[:, i:].t(), w_weight[:, :i].t(), w_weight[:, i:].t(), zp], 1).contiguous()
        zc, zh = torch.zeros(C_, device=x.device), torch.zeros(H, device=x.device)
        bcat = torch.cat([zc, f_bias if f_bias is not None else zc, zh, w_bias if w_bias is not None else zh, zp[0]])
        P = linear(x, wcat, bcat)
        out = torch.empty(graph.n_nodes, C_, dtype=torch.float32, device=x.device)
        den = torch.empty(graph.n_nodes, H, dtype=torch.float32, device=x.device)
        amax = torch.empty(H, dtype=torch.int64, device=x.device)
        nb = lib.gode_gat_fwd_workspace_bytes(C.byref(graph.c), H, oh)
        ws = workspace(nb, x.device, "gat_fwd")
        check(lib.gode_gat_fwd(C.byref(graph.c), H, oh, _p(P), ldp, float(eps), _p(out), out.stride(0), _p(den), _p(amax),
                               _p(graph.nan_flag), _p(ws), nb, _stream()), "gode_gat_fwd")
        ctx.graph, ctx.H, ctx.oh, ctx.i = graph, H, oh, i
        ctx.has_bias = (f_bias is not None, w_bias is not None)
        ctx.save_for_backward(x, wcat, P, out, den, amax)
        return out

    @staticmethod
    def backward(ctx, g):
        x, wcat, P, out, den, amax = ctx.saved_tensors
        graph, H, oh, i = ctx.graph, ctx.H, ctx.oh, ctx.i
        C_ = H * oh
        ldp = P.shape[1]
        g = _rowmajor(g, "grad").contiguous()
        dP = torch.empty_like(P)
        if ldp != 2 * C_ + 2 * H:
            dP[:, 2 * C_ + 2 * H:].zero_()
        nb = lib.gode_gat_bwd_workspace_bytes(C.byref(graph.c), H, oh)
        ws = workspace(nb, x.device, "gat")
        check(lib.gode_gat_bwd(C.byref(graph.c), H, oh, _p(P), ldp, _p(out), out.stride(0), _p(den), _p(amax), _p(g),
                               g.stride(0), _p(dP), _p(ws), nb, _stream()), "gode_gat_bwd")
        big = _tc_gemm_enabled(x.shape[0], min(i, ldp), min(i, ldp))
        gx = None
        if ctx.needs_input_grad[0]:
            gx = gemm_tc(dP, wcat) if big else gemm(dP, wcat, trans_b=True)              # dP [N, ldp] x wcat [i, ldp]^T
        gfw = gfb = gww = gwb = None
        if any(ctx.needs_input_grad[1:5]):
            # x^T dP reduces over the N rows of two tall, narrow operands: the SIMT split-K kernel reads them as they lie
            # (5.7 ms at N = 1 M); transposing both for the K-major tensor-core kernel costs more than it saves here
            dw = gemm(x, dP, trans_a=True, splits=_splits_for(x.shape[0], i, ldp))       # [i, ldp]
            gfw = torch.cat([dw[:, :C_].t(), dw[:, C_:2 * C_].t()], 1)
            gww = torch.cat([dw[:, 2 * C_:2 * C_ + H].t(), dw[:, 2 * C_ + H:2 * C_ + 2 * H].t()], 1)
            db = colsum(dP)
            gfb = db[C_:2 * C_] if ctx.has_bias[0] else None
            gwb = db[2 * C_ + H:2 * C_ + 2 * H] if ctx.has_bias[1] else None
        return gx, gfw, gfb, gww, gwb, None, None, None


def gn_wgrad(y, groups, eps, g):
    """``(xhat(y)^T g, colsum(g))`` with xhat = GroupNorm(y) before its affine -- gode_gn_wgrad_f32 (tcgen05)."""
    y, g = _rowmajor(y, "y"), _rowmajor(g, "g")
    d, nc = y.shape[1], g.shape[1]
    out = torch.empty(d, nc, dtype=torch.float32, device=y.device)
    cs = torch.empty(nc, dtype=torch.float32, device=y.device)
    nb = lib.gode_gn_wgrad_workspace_bytes(d)
    ws = workspace(nb, y.device, "gn_wgrad")
    check(lib.gode_gn_wgrad_f32(y.shape[0], d, groups, float(eps), _p(y), _p(g), g.stride(0), nc, _p(out), nc, _p(cs), _p(ws), nb,
                                _lib.PREC_FP32, _stream()), "gode_gn_wgrad_f32")
    return out, cs


def gat_ode_fusable(y, groups, heads, oh):
    """The fused GAT ODE function covers the widths its tensor-core weight gradient does (d = 128, 32 groups)."""
    import os
    return (os.environ.get("GODE_GAT_FUSED", "1") != "0" and y.is_cuda and y.dim() == 2 and y.shape[1] == 128 and groups == 32
            and y.shape[0] >= 1)


class GatOdeFn(torch.autograd.Function):
    """f(t, y) = GATconv([t*1 || GroupNorm(y)])  (GAT/models.py:161-179 over GAT/layers.py:40-58) as ONE autograd node.

    The t column never materialises: with W_cat = [Wf_src^T | Wf_tgt^T | ww_src^T | ww_tgt^T] ([d+1, ldp], row 0 the t row)
    the node projection is  P = GroupNorm(y) W_cat[1:] + (t W_cat[0] + b_cat),  a K = d = 128 product on aligned rows (the
    concatenated [N, 129] operand has 516-byte rows: no 128-bit access, a padded copy per call and a second, 99 % empty
    column tile in the input-gradient product).  Backward: gode_gat_bwd -> dP; input gradient dP W_cat[1:]^T -> GroupNorm
    backward; weight gradient xhat(y)^T dP on tcgen05 (gode_gn_wgrad_f32) with the GroupNorm affine, the t row and both
    biases recovered from the column sums of dP.
    """

    @staticmethod
    def forward(ctx, y, t, gamma, beta, f_weight, f_bias, w_weight, w_bias, graph, heads, eps, groups, gn_eps):
        y = _rowmajor(y, "y")
        d = y.shape[1]
        i = d + 1
        C_, H = f_weight.shape[0], heads
        oh = C_ // H
        ldp = 2 * C_ + 2 * H
        pad = (-ldp) % 4        # rows of P are padded to a multiple of 16 bytes (one head: 2 * oh + 2 columns) with zero weights
        # W1 [d, ldp + pad] = rows 1.. of W_cat ([in, out] layout), the t columns (0 and i) of the Linear weights dropped
        zp = torch.zeros(d, pad, device=y.device)
        w1 = torch.cat([f_weight[:, 1:i].t(), f_weight[:, i + 1:].t(), w_weight[:, 1:i].t(), w_weight[:, i + 1:].t(), zp], 1)
        w0 = torch.cat([f_weight[:, 0], f_weight[:, i], w_weight[:, 0], w_weight[:, i], zp[0]])      # the t row of W_cat
        zc, zh = torch.zeros(C_, device=y.device), torch.zeros(H, device=y.device)
        bcat = torch.cat([zc, f_bias if f_bias is not None else zc, zh, w_bias if w_bias is not None else zh, zp[0]])
        ldp += pad
        tt = t.detach().to(torch.float32) if torch.is_tensor(t) else float(t)
        row0 = w0 * tt + bcat                      # stays on the device: t is a device scalar inside the adjoint solve
        if lib.gode_gn_linear_supported(d, groups, ldp):
            # GroupNorm inside the producer of the tcgen05 product; [row0; W1] is the [d + 1, ldp] weight with "t" = 1
            wfull = torch.cat([row0[None], w1], 0).contiguous()
            P = torch.empty(y.shape[0], ldp, dtype=torch.float32, device=y.device)
            check(lib.gode_gn_linear_f32(y.shape[0], d, groups, float(gn_eps), _p(y), _p(gamma.contiguous()), _p(beta.contiguous()),
                                         _p(wfull), ldp, 1.0, ldp, _p(P), ldp, _lib.PREC_FP32, _stream()), "gode_gn_linear_f32")
        else:
            xn = groupnorm_fwd(y, groups, gamma, beta, gn_eps)
            P = gemm_tc(xn, w1.t().contiguous(), row0)
            del xn
        w1 = w1.contiguous()
        out = torch.empty(graph.n_nodes, C_, dtype=torch.float32, device=y.device)
        den = torch.empty(graph.n_nodes, H, dtype=torch.float32, device=y.device)
        amax = torch.empty(H, dtype=torch.int64, device=y.device)
        nb = lib.gode_gat_fwd_workspace_bytes(C.byref(graph.c), H, oh)
        ws = workspace(nb, y.device, "gat_fwd")
        check(lib.gode_gat_fwd(C.byref(graph.c), H, oh, _p(P), ldp, float(eps), _p(out), out.stride(0), _p(den), _p(amax),
                               _p(graph.nan_flag), _p(ws), nb, _stream()), "gode_gat_fwd")
        ctx.graph, ctx.H, ctx.oh, ctx.groups, ctx.gn_eps = graph, H, oh, groups, gn_eps
        ctx.has_bias = (f_bias is not None, w_bias is not None)
        ctx.t = tt
        ctx.save_for_backward(y, gamma, beta, w1, w0, P, out, den, amax)
        return out

    @staticmethod
    def backward(ctx, g):
        y, gamma, beta, w1, w0, P, out, den, amax = ctx.saved_tensors
        graph, H, oh = ctx.graph, ctx.H, ctx.oh
        C_ = H * oh
        ldp = P.shape[1]                # 2 * C_ + 2 * H, padded to a multiple of 4
        g = _rowmajor(g, "grad").contiguous()
        dP = torch.empty_like(P)
        if ldp != 2 * C_ + 2 * H:
            dP[:, 2 * C_ + 2 * H:].zero_()
        nb = lib.gode_gat_bwd_workspace_bytes(C.byref(graph.c), H, oh)
        ws = workspace(nb, y.device, "gat")
        check(lib.gode_gat_bwd(C.byref(graph.c), H, oh, _p(P), ldp, _p(out), out.stride(0), _p(den), _p(amax), _p(g),
                               g.stride(0), _p(dP), _p(ws), nb, _stream()), "gode_gat_bwd")
        need = ctx.needs_input_grad
        gy = gt = ggam = gbeta = gfw = gfb = gww = gwb = None
        if need[0] or need[2] or need[3]:
            dxn = gemm_tc(dP, w1)                                  # dP [N, ldp] x W1^T [ldp, d]: one 128-column tile
            gy, ggam, gbeta = groupnorm_bwd(y, ctx.groups, gamma, dxn, ctx.gn_eps)
            del dxn
        if any(need[4:8]) or need[1]:
            raw, cs = gn_wgrad(y, ctx.groups, ctx.gn_eps, dP)      # xhat^T dP [d, ldp], colsum(dP) [ldp]
            if need[1]:
                gt = (cs * w0).sum()
                if torch.is_tensor(ctx.t):
                    gt = gt.reshape(ctx.t.shape)
            if any(need[4:8]):
                dw1 = gamma[:, None] * raw + beta[:, None] * cs[None, :]       # rows 1..d of dW_cat
                dw0 = cs * ctx.t                                               # the t row
                def half(a, b):   # columns [a, b) of dW_cat, as [b - a, d + 1] (nn.Linear layout: [out, in])
                    return torch.cat([dw0[a:b, None], dw1[:, a:b].t()], 1)
                gfw = torch.cat([half(0, C_), half(C_, 2 * C_)], 1)
                gww = torch.cat([half(2 * C_, 2 * C_ + H), half(2 * C_ + H, 2 * C_ + 2 * H)], 1)
                gfb = cs[C_:2 * C_] if ctx.has_bias[0] else None
                gwb = cs[2 * C_ + H:2 * C_ + 2 * H] if ctx.has_bias[1] else None
        return gy, gt, ggam, gbeta, gfw, gfb, gww, gwb, None, None, None, None, None


def gat_ode_func(y, t, norm, conv, check_nan=True):
    """Fused evaluation of the GAT ODE function for ``norm`` (nn.GroupNorm) and ``conv`` (GAT FixedGraphConvolution)."""
    graph = gat_graph_for(conv.src, conv.tgt, y.shape[0])
    out = GatOdeFn.apply(y, t, norm.weight, norm.bias, conv.f.weight, conv.f.bias, conv.w.weight, conv.w.bias, graph,
                         conv.heads, conv.eps, norm.num_groups, norm.eps)
    if check_nan:
        assert int(graph.nan_flag.item()) == 0, "NaN in GAT attention"
    return out


def gat_conv(x, src, tgt, f_weight, f_bias, w_weight, w_bias, heads=1, eps=1e-6, check_nan=True):
    graph = gat_graph_for(src, tgt, x.shape[0])
    out = GatConvFn.apply(x, f_weight, f_bias, w_weight, w_bias, graph, heads, eps)
    if check_nan:
        # the reference asserts on NaNs six times per call (GAT/layers.py:46-56); one device flag, one read
        assert int(graph.nan_flag.item()) == 0, "NaN in GAT attention"
    return out


# --------------------------------------------------------------------------------------------------
# segmented sums through a plan (scatter_add readout, one-hot incidence products) and QC edge messages
# --------------------------------------------------------------------------------------------------


class PlanMatmulFn(torch.autograd.Function):
    """``A @ x`` for a planned sparse A (forward gode_spmm_csr_f32, backward the same kernel on A^T)."""

    @staticmethod
    def forward(ctx, x, plan, bias):
        ctx.plan, ctx.has_bias = plan, bias is not None
        return spmm(plan, x, bias=bias)

    @staticmethod
    def backward(ctx, g):
        g = _rowmajor(g, "grad")
        gx = spmm(ctx.plan, g, transpose=True) if ctx.needs_input_grad[0] else None
        gb = colsum(g) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, None, gb


def index_plan(index, n_rows):
    """Plan of the one-hot matrix M[index[j], j] = 1 (``scatter_add(x, index)`` == ``M @ x``), cached per index tensor."""
    key = ("idx", index.data_ptr(), int(index.numel()), int(n_rows), index.device.index)
    hit = _plan_cache.get(key)
    if hit is not None and hit[0]() is index:
        return hit[1]
    if not index.is_cuda:
        raise TypeError("index must live on a CUDA device; graph-odenet_b200 has no CPU path")
    j = torch.arange(index.numel(), device=index.device, dtype=torch.int64)
    plan = GraphPlan.from_coo(index.to(torch.int64), j, torch.ones(index.numel(), dtype=torch.float32, device=index.device),
                              int(n_rows), int(index.numel()))
    _plan_cache[key] = (weakref.ref(index), plan)
    return plan


def scatter_add_rows(x, index, n_rows):
    """``torch_scatter.scatter_add(x, index, dim=0, dim_size=n_rows)`` (QC/layer_models.py:121, QC/torch_scatter.py)
    as a deterministic segmented sum."""
    return PlanMatmulFn.apply(x, index_plan(index, n_rows), None)


class SegmentAttendFn(torch.autograd.Function):
    """r[g] = sum_i softmax_g(<x_i, q_g>) x_i over the nodes of graph g (QC/set2set.py:60-75) -- gode_segment_attend_*."""

    @staticmethod
    def forward(ctx, x, q, gptr):
        x, q = _rowmajor(x, "x"), _rowmajor(q, "q")
        B, h = q.shape
        a = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        r = torch.empty(B, h, dtype=torch.float32, device=x.device)
        check(lib.gode_segment_attend_fwd(B, h, _p(gptr), _p(x), x.stride(0), _p(q), q.stride(0), _p(a), _p(r), r.stride(0),
                                          _stream()), "gode_segment_attend_fwd")
        ctx.save_for_backward(x, q, a, gptr)
        return r

    @staticmethod
    def backward(ctx, dr):
        x, q, a, gptr = ctx.saved_tensors
        dr = _rowmajor(dr, "grad").contiguous()
        B, h = q.shape
        dx = torch.empty_like(x) if x.is_contiguous() else torch.empty(x.shape, dtype=torch.float32, device=x.device)
        dq = torch.empty(B, h, dtype=torch.float32, device=x.device)
        check(lib.gode_segment_attend_bwd(B, h, _p(gptr), _p(x), x.stride(0), _p(q), q.stride(0), _p(a), _p(dr), dr.stride(0),
                                          _p(dx), dx.stride(0), _p(dq), dq.stride(0), _stream()), "gode_segment_attend_bwd")
        return dx, dq, None


def segment_attend(x, q, gptr):
    """``gptr`` int32 [B + 1]: nodes of graph g are rows gptr[g] .. gptr[g+1] of ``x``."""
    return SegmentAttendFn.apply(x, q, gptr)


def incidence_plan(Etgt, n_nodes=None):
    """Plan of the reference's ``Etgt`` (dense one-hot [N, E], QC/datasets/utils.py:214; also accepted: sparse COO,
    or the target index vector [E] with ``n_nodes``)."""
    if isinstance(Etgt, GraphPlan):
        return Etgt
    if Etgt.dim() == 1:
        return index_plan(Etgt, n_nodes)
    return plan_for(Etgt)


class EdgeMessageFn(torch.autograd.Function):
    """out = Etgt @ (edge_data[e] @ s[Esrc[e]])_e (+ bias)  -- QC/layers.py:143-147, QC/mpnn.py:27-29."""

    @staticmethod
    def forward(ctx, s, edge_data, esrc, plan_t, bias):
        s = _rowmajor(s, "support")
        ed = _req(edge_data, "edge_data").contiguous()
        E, f = ed.shape[0], ed.shape[1]
        if ed.dim() != 3 or ed.shape[2] != f or s.shape[1] != f:
            raise ValueError("edge_data must be [E, f, f] with f = support width")
        esrc32 = esrc.to(torch.int32).contiguous()
        if E == 0:      # no edges: the aggregate is the bias alone
            out = torch.zeros(s.shape[0], f, dtype=torch.float32, device=s.device)
            ctx.plan_t, ctx.has_bias, ctx.n = plan_t, bias is not None, s.shape[0]
            ctx.save_for_backward(s, ed, esrc32, esrc)
            return out if bias is None else out + bias
        msg = torch.empty(E, f, dtype=torch.float32, device=s.device)
        check(lib.gode_edge_matvec(E, f, _p(ed), _p(esrc32), _p(s), s.stride(0), _p(msg), _stream()), "gode_edge_matvec")
        out = spmm(plan_t, msg, bias=bias)
        ctx.plan_t, ctx.has_bias, ctx.n = plan_t, bias is not None, s.shape[0]
        ctx.save_for_backward(s, ed, esrc32, esrc)
        return out

    @staticmethod
    def backward(ctx, g):
        s, ed, esrc32, esrc = ctx.saved_tensors
        g = _rowmajor(g, "grad")
        E, f = ed.shape[0], ed.shape[1]
        if E == 0:
            return (torch.zeros_like(s), torch.zeros_like(ed), None, None,
                    colsum(g) if (ctx.has_bias and ctx.needs_input_grad[4]) else None)
        dm = spmm(ctx.plan_t, g, transpose=True)                    # Etgt^T g : [E, f]
        ds_msg = torch.empty(E, f, dtype=torch.float32, device=g.device)
        d_ed = torch.empty_like(ed) if ctx.needs_input_grad[1] else None
        check(lib.gode_edge_matvec_bwd(E, f, _p(ed), _p(esrc32), None, _p(s), s.stride(0), _p(dm), dm.stride(0),
                                       _p(ds_msg), _p(d_ed), _stream()), "gode_edge_matvec_bwd")
        ds = spmm(index_plan(esrc, ctx.n), ds_msg) if ctx.needs_input_grad[0] else None
        gb = colsum(g) if (ctx.has_bias and ctx.needs_input_grad[4]) else None
        return ds, d_ed, None, None, gb


def edge_message(s, edge_data, esrc, Etgt, bias=None):
    plan_t = incidence_plan(Etgt, s.shape[0])
    if plan_t.n_rows != s.shape[0] or plan_t.n_cols < edge_data.shape[0]:
        raise ValueError("Etgt must be [N, E]")
    return EdgeMessageFn.apply(s, edge_data, esrc, plan_t, bias)


# --------------------------------------------------------------------------------------------------
# trainer epilogue: log_softmax + masked nll_loss + accuracy, and Adam, as single libgode launches
# --------------------------------------------------------------------------------------------------


def index_mask(idx, n):
    """uint8 [n] with ones at ``idx`` (built once per split; the fused loss takes the index set as a mask)."""
    m = torch.zeros(n, dtype=torch.uint8, device=idx.device)
    m[idx] = 1
    return m


class LogSoftmaxNllFn(torch.autograd.Function):
    """(logp, [loss, acc]) = gode_lsm_nll_fwd(z): ``F.log_softmax(z, 1)``, ``F.nll_loss(logp[idx], labels[idx])`` and
    ``accuracy(logp[idx], labels[idx])`` (GCN/train_res.py:76-77, GCN/utils.py:215-219) in one pass; the backward of the loss
    is one more launch.  ``logp`` is returned for the caller's use but carries no gradient path of its own."""

    @staticmethod
    def forward(ctx, z, labels, mask, m):
        z = _rowmajor(z, "logits")
        n, c = z.shape
        logp = torch.empty(n, c, dtype=torch.float32, device=z.device)
        la = torch.empty(2, dtype=torch.float32, device=z.device)
        nb = lib.gode_lsm_nll_workspace_bytes(n)
        ws = workspace(nb, z.device, "lsm")
        check(lib.gode_lsm_nll_fwd(n, c, _p(z), z.stride(0), _p(labels), _p(mask), int(m), _p(logp), c, _p(la), _p(ws), nb,
                                   _stream()), "gode_lsm_nll_fwd")
        ctx.m = int(m)
        ctx.save_for_backward(logp, labels, mask)
        ctx.mark_non_differentiable(logp)
        return logp, la

    @staticmethod
    def backward(ctx, _glogp, gla):
        logp, labels, mask = ctx.saved_tensors
        n, c = logp.shape
        gla = gla.contiguous()
        dz = torch.empty(n, c, dtype=torch.float32, device=logp.device)
        check(lib.gode_lsm_nll_bwd(n, c, _p(logp), c, _p(labels), _p(mask), ctx.m, _p(gla), _p(dz), c, _stream()),
              "gode_lsm_nll_bwd")
        return dz, None, None, None


def log_softmax_nll(z, labels, mask, m):
    """Returns ``(logp [N, C], loss_acc [2])`` with ``loss_acc[0]`` the mean NLL over the masked rows (differentiable) and
    ``loss_acc[1]`` their accuracy."""
    if labels.dtype != torch.int64 or mask.dtype != torch.uint8:
        raise TypeError("labels must be int64 class ids and mask uint8")
    return LogSoftmaxNllFn.apply(z, labels.contiguous(), mask.contiguous(), m)


class FusedAdam:
    """``torch.optim.Adam(params, lr, betas, eps, weight_decay)`` (GCN/train_res.py:126-127) as ONE libgode launch per step:
    the parameters are re-pointed at slices of one flat buffer, so the update runs over [sum numel] elements at once; the
    step counter lives on the device (a replayed CUDA graph keeps counting).  Same update rule and defaults as torch's."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise TypeError("FusedAdam needs CUDA parameters; graph-odenet_b200 has no CPU path")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + k].view(p.shape)
                off += k
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.steps = torch.zeros(1, dtype=torch.int32, device=dev)

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def step(self):
        g = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params])
        check(lib.gode_adam_step(self.flat.numel(), _p(self.flat), _p(g), _p(self.m), _p(self.v), self.lr, self.betas[0],
                                 self.betas[1], self.eps, self.weight_decay, _p(self.steps), _stream()), "gode_adam_step")
