"""Row-partitioned (multi-GPU) GCN-ODE: 1-D contiguous row blocks of A_hat and of every [N, d] tensor, one
process per GPU, a halo exchange of the gather operand per function evaluation.

The reference is single-device (no torch.distributed anywhere); this is the sharding BASELINE.json's
config 4 asks for.  What is partitioned is exactly the reference's ``torch.spmm(self.adj, support)``
(GCN/layers.py:71) and its autograd transpose: rank p owns rows [lo_p, hi_p) of A_hat (forward gather) and
of A_hat^T (backward gather).  Columns are renumbered to ``[owned | halo]``, halo = the sorted remote row ids the
block references, so the local kernels (``gode_gcn_*``) run unchanged on an operand buffer with a halo tail.

Per forward evaluation: transform (row-local) -> pack the rows peers reference (``gode_gather_rows``) ->
``all_to_all_single`` straight into the halo tail of S -> gather.  Per adjoint evaluation a second exchange
moves gP for the A_hat^T gather.  Parameter gradients are per-rank partial sums, summed once per backward
(fixed-step methods) or once per step (adaptive methods, whose error norm needs the global value).
GroupNorm, the dense products and the Runge-Kutta arithmetic are row-local.

Index work (partition bounds, halo lists, renumbering) is exact integer arithmetic; the same code runs on CPU
tensors with the ``gloo`` backend, which is how tests/test_parallel_cpu.py checks it without a GPU.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import ops
from ._lib import lib, check


def partition_bounds(n, world):
    """Row bounds [b_0=0, ..., b_world=n] of the contiguous 1-D partition (sizes differ by at most one)."""
    return [(n * r) // world for r in range(world + 1)]


def _owner_counts(ids, bounds):
    """How many of the sorted ``ids`` fall in each rank's range."""
    b = torch.as_tensor(bounds, dtype=torch.int64, device=ids.device)
    pos = torch.searchsorted(ids, b)          # first index with id >= bound
    return (pos[1:] - pos[:-1]).tolist()


class HaloPlan:
    """Who sends which rows to whom for one gather operand.

    ``halo`` (int64, sorted, global ids) are the remote rows this rank reads; because the partition is contiguous
    the sorted order is also ordered by owner, so the receive side of the all-to-all is the halo tail as is.
    ``send_idx`` (int32, local row ids, ordered by destination then id) is what this rank packs for its peers.
    """

    def __init__(self, halo, bounds, rank, group=None):
        self.group, self.rank = group, rank
        self.world = len(bounds) - 1
        self.lo, self.hi = bounds[rank], bounds[rank + 1]
        self.n_own = self.hi - self.lo
        self.halo = halo
        self.n_halo = int(halo.numel())
        dev = halo.device
        self.recv_counts = _owner_counts(halo, bounds)
        self._send_buf = None
        assert self.recv_counts[rank] == 0, "halo contains owned rows"
        if self.world == 1:
            self.send_counts, self.send_idx = [0], torch.empty(0, dtype=torch.int32, device=dev)
            return
        rc = torch.tensor(self.recv_counts, dtype=torch.int64, device=dev)
        sc = torch.empty_like(rc)
        dist.all_to_all_single(sc, rc, group=group)
        self.send_counts = sc.tolist()
        ids = torch.empty(sum(self.send_counts), dtype=torch.int64, device=dev)
        dist.all_to_all_single(ids, halo.contiguous(), self.send_counts, self.recv_counts, group=group)
        if ids.numel() and (int(ids.min()) < self.lo or int(ids.max()) >= self.hi):
            raise IndexError("a peer requested rows this rank does not own")
        self.send_idx = (ids - self.lo).to(torch.int32)

    def bytes_per_exchange(self, d):
        return (sum(self.send_counts) + self.n_halo) * d * 4

    def exchange(self, buf, pack=None):
        """Fill ``buf[n_own:]`` (the halo tail) with the peers' rows; ``buf[:n_own]`` must hold the owned rows.
        ``pack(buf, idx) -> rows`` replaces the CUDA pack kernel (the gloo/CPU tests pass an index_select)."""
        if self.world == 1:
            return buf
        d = buf.shape[1]
        n_send = int(self.send_idx.numel())
        if pack is not None:
            send = pack(buf, self.send_idx)
        else:
            ops._req(buf, "halo operand")
            if self._send_buf is None or self._send_buf.shape != (n_send, d) or self._send_buf.device != buf.device:
                self._send_buf = torch.empty(n_send, d, dtype=torch.float32, device=buf.device)
            send = self._send_buf
            check(lib.gode_gather_rows(n_send, ops._p(self.send_idx), d, ops._p(buf), buf.stride(0), ops._p(send), d,
                                       ops._stream()), "gode_gather_rows")
        rows = lambda k: [int(c) for c in k]
        dist.all_to_all_single(buf[self.n_own:], send, rows(self.recv_counts), rows(self.send_counts), group=self.group)
        return buf


def local_block(row, col, val, bounds, rank, transpose=False):
    """Entries of A (or A^T) whose row this rank owns, renumbered to local ids.

    Returns (r_local, c_local, v, halo): r_local in [0, n_own), c_local in [0, n_own + n_halo) with owned columns
    first (col - lo) and remote columns mapped to n_own + (position in the sorted halo list).  Pure index
    arithmetic -- bit-exact by construction; un-renumbering (``global_cols``) gives back the original ids.
    """
    if transpose:
        row, col = col, row
    lo, hi = bounds[rank], bounds[rank + 1]
    m = (row >= lo) & (row < hi)
    r, c, v = row[m] - lo, col[m], val[m]
    remote = (c < lo) | (c >= hi)
    halo = torch.unique(c[remote])
    n_own = hi - lo
    c_local = torch.where(remote, n_own + torch.searchsorted(halo, c), c - lo)
    return r, c_local, v, halo


def global_cols(c_local, halo, lo, n_own):
    """Inverse of the column renumbering of ``local_block``."""
    own = c_local < n_own
    return torch.where(own, c_local + lo, halo[(c_local - n_own).clamp_(min=0)])


class PartitionedPlan:
    """This rank's row block of A_hat and of A_hat^T with their halo plans (what ``GcnKernel`` needs from a plan)."""

    def __init__(self, n, bounds, rank, A, At, halo, halo_t, nnz_global, group=None):
        self.n_global, self.bounds, self.rank, self.group = n, bounds, rank, group
        self.world = len(bounds) - 1
        self.lo, self.hi = bounds[rank], bounds[rank + 1]
        self.n_rows = self.hi - self.lo
        self.A, self.At, self.halo, self.halo_t = A, At, halo, halo_t
        self.nnz = A.nnz
        self.nnz_global = nnz_global
        self.rowptr_t = At.rowptr            # "has a transpose" marker GcnKernel looks at
        self.device = A.device

    @classmethod
    def build(cls, row, col, val, n, rank=None, world=None, group=None, bounds=None):
        """``row, col, val``: the COO of A_hat (int64, int64, fp32) -- at least every entry whose row or column
        this rank owns; the full COO is fine."""
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        bounds = bounds or partition_bounds(n, world)
        n_own = bounds[rank + 1] - bounds[rank]
        r, c, v, halo = local_block(row, col, val, bounds, rank)
        A = ops.GraphPlan.from_coo(r, c, v, n_own, n_own + int(halo.numel()), build_transpose=False)
        del r, c, v
        r, c, v, halo_t = local_block(row, col, val, bounds, rank, transpose=True)
        At = ops.GraphPlan.from_coo(r, c, v, n_own, n_own + int(halo_t.numel()), build_transpose=False)
        del r, c, v
        return cls(n, bounds, rank, A, At, HaloPlan(halo, bounds, rank, group), HaloPlan(halo_t, bounds, rank, group),
                   int(val.numel()), group)

    def csr(self, transpose=False):
        return (self.At if transpose else self.A).csr(False)

    def make_kernel(self, *args):
        from .odeint import GcnKernel

        plan = self

        class PartitionedGcnKernel(GcnKernel):
            """GcnKernel on a row block: halo exchanges after every producer of a gather operand."""

            def new_S(self):
                return torch.empty(plan.n_rows + plan.halo.n_halo, self.d, dtype=torch.float32, device=self.dev)

            def new_gP(self):
                return torch.empty(plan.n_rows + plan.halo_t.n_halo, self.d, dtype=torch.float32, device=self.dev)

            def numel_global(self):
                return plan.n_global * self.d

            def scalar(self, dev_scalar):
                if plan.world > 1:
                    dist.all_reduce(dev_scalar, group=plan.group)
                return float(dev_scalar.item())

            def reduce_small(self, t):
                if plan.world > 1:
                    if t.is_contiguous():
                        dist.all_reduce(t, group=plan.group)
                    else:
                        c = t.contiguous()
                        dist.all_reduce(c, group=plan.group)
                        t.copy_(c)
                return t

            def transform(self, y, t, out):
                super().transform(y, t, out)
                plan.halo.exchange(out)
                return out

            def stage_fwd(self, S, k_out, y0=None, kprev=(), coefs=(), coef_self=0.0, y_next=None, t_next=0.0,
                          S_next=None):
                super().stage_fwd(S, k_out, y0, kprev, coefs, coef_self, y_next, t_next, S_next)
                if S_next is not None:
                    plan.halo.exchange(S_next)

            def vjp_phase1(self, S, a, sign, k_y, gP, y0=None, kprev=(), coefs=(), coef_self=0.0, y_next=None):
                super().vjp_phase1(S, a, sign, k_y, gP, y0, kprev, coefs, coef_self, y_next)
                plan.halo_t.exchange(gP)

        return PartitionedGcnKernel(self, *args)

    def halo_bytes_per_step(self, d, n_fwd_evals, n_aug_evals):
        """Bytes this rank moves over NVLink (sent + received) in one fwd+bwd step."""
        s = self.halo.bytes_per_exchange(d)
        g = self.halo_t.bytes_per_exchange(d)
        return (n_fwd_evals + n_aug_evals) * s + n_aug_evals * g
