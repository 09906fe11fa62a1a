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
from ._lib import lib, check, MAX_STAGES


def init_process_group(device, backend="nccl"):
    """One process per GPU (RANK / WORLD_SIZE / MASTER_* from the launcher).  The NCCL stream is created with high
    priority so that a halo exchange issued underneath a gather or a dense kernel is scheduled as soon as SM
    resources free up instead of behind the compute kernel's remaining CTAs."""
    import os
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    kw = {}
    if backend == "nccl":
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        kw = dict(pg_options=opts, device_id=device)
    dist.init_process_group(backend, **kw)


MODES = ("sync", "async", "split", "p2p", "p2p-async", "p2p-fused")


def pipe_g_chunks(world):
    """Row chunks of the pipelined forward stage (GODE_PIPE_G; default 0 = off).  Measured at 2 GPUs (N = 10 M,
    profiles/r02_multi_gpu.md): 113.0 ms per step without, 112.9 with 2 chunks, 114.1 with 4 -- the exposed exchange time
    drops by 4 ms and the gathers the pushes now run next to slow down by as much, the same trade GODE_PIPE_S showed in
    round 1.  Parity-tested (tests/test_gpu_parallel.py, modes p2p-fused:g2 / :g3), opt-in."""
    import os
    return int(os.environ.get("GODE_PIPE_G", "0"))


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


def split_block(r, c_local, v, n_own):
    """Column split of a renumbered block: (owned-column entries, halo-column entries with columns re-based to 0)."""
    own = c_local < n_own
    halo = ~own
    return (r[own], c_local[own], v[own]), (r[halo], c_local[halo] - n_own, v[halo])


class HaloKernelMixin:
    """GcnKernel on a row block: a halo exchange after every producer of a gather operand, awaited by the kernel that
    gathers from the buffer.  Mixed in front of odeint.GcnKernel (``plan.make_kernel``); kept free of CUDA calls of its
    own so that tests/test_peer_protocol_cpu.py can replay the solver's control flow over a recording base class."""

    def __init__(self, plan, *a):
        super().__init__(plan, *a)
        self.pending = {}            # data_ptr of an operand buffer -> event / epoch of its in-flight halo exchange
        self.partial = None
        self.peer = plan.peer_for(self.d)
        self.fused = self.peer is not None and plan.mode == "p2p-fused" and self._push_fusable()
        # The support S is pushed by the stand-alone kernel unless GODE_FUSE_S=1: remote stores need many warps in flight
        # to reach NVLink speed, and the warp-specialised transform has four epilogue warps per SM (measured at 2 GPUs,
        # N = 10 M: 4.2 ms per transform with the push in its epilogue vs 1.2 ms + 1.2 ms for transform + push kernel);
        # the phase-1 gather, with 64 resident warps per SM, hides its gP push completely and keeps it fused.
        import os
        self.fused_S = self.fused and os.environ.get("GODE_FUSE_S", "0") == "1"
        # GODE_PIPE_S = number of row chunks (default 0 = off): the transform and the push of S pipelined over row chunks --
        # chunk c is pushed on the side stream while chunk c+1 is transformed, and in the adjoint the whole exchange rides
        # underneath phase 2; all flag writes then go to the side stream.  Measured at 2 GPUs (N = 10 M): 119.1 ms per step
        # with 4 chunks vs 119.9 ms without -- what the overlap hides (outside 16.8 -> 8.3 ms) the concurrent push takes back
        # from the kernels it runs next to (transform 9.3 -> 14.5, A^T gather 18.8 -> 20.5 ms) -- so it stays opt-in.
        self.pipe_S = int(os.environ.get("GODE_PIPE_S", "0")) if (self.fused and not self.fused_S and self.peer.side) else 0
        # GODE_PIPE_G = number of row chunks (default 0 = off, see pipe_g_chunks): a forward stage is GATHERED in row chunks -- chunk c's rows of y_next are transformed and pushed on the side
        # stream while chunk c+1 is still being gathered on the main stream (posted NVLink stores next to a kernel with 64
        # resident warps per SM, as the fused gP push), so only the last chunk's push stays exposed.  Needs the tile gather
        # (d = 128) and the side stream.
        self.pipe_G = (pipe_g_chunks(plan.world)
                       if (self.fused and not self.fused_S and not self.pipe_S and self.peer.side and self.d == 128
                           and plan.split is None) else 0)
        # GODE_PUSH_Y (default 1 with the fused exchange): the halo exchange moves from the transform's OUTPUT to its INPUT.
        # The gather that produces a stage state y_next stores the rows its peers reference into their copies of the
        # buffer (gode_gcn_odefunc_t.push_y: posted NVLink writes underneath a kernel with 64 resident warps per SM, as
        # the gP push) and every rank transforms [owned | halo] rows itself -- the owned rows while the peers' rows are
        # still landing.  The stand-alone push of S (8 per rk4 step, nothing to hide under in the forward chain
        # gather -> transform -> push -> gather) and its exposed wait disappear; the price is the transform of the halo rows.
        self.push_y = (self.fused and not self.fused_S and not self.pipe_S and not self.pipe_G and plan.split is None
                       and os.environ.get("GODE_PUSH_Y", "1") != "0")
        if plan.split is not None:
            from . import _lib
            # descriptor for the second (halo-column) pass: same parameters, halo blocks, operand offset
            h = _lib.GcnOdeFunc()
            C.memmove(C.byref(h), C.byref(self.f), C.sizeof(h))
            h.A = plan.split["A_halo"].csr(False)
            h.At = plan.split["At_halo"].csr(False)
            h.gather_row_offset = plan.n_rows
            self.f_halo = h
            self.csr_own = plan.split["A_own"].csr(False)
            self.csr_own_t = plan.split["At_own"].csr(False)
            self.ws_bytes = max(self.ws_bytes, lib.gode_gcn_workspace_bytes(C.byref(h)))

    def new_S(self):
        plan = self.plan
        if self.peer is not None:
            return self.peer.new(plan.n_rows + plan.halo.n_halo)
        return torch.empty(plan.n_rows + plan.halo.n_halo, self.d, dtype=torch.float32, device=self.dev)

    def new_gP(self):
        plan = self.plan
        if self.peer is not None:
            return self.peer.new(plan.n_rows + plan.halo_t.n_halo)
        return torch.empty(plan.n_rows + plan.halo_t.n_halo, self.d, dtype=torch.float32, device=self.dev)

    def new_Y(self):
        if not self.push_y:
            return self.new()
        plan = self.plan
        full = self.peer.new(plan.n_rows + plan.halo.n_halo)
        view = self._own_rows(full)
        if view is full:
            full._gode_y = True          # (the protocol simulator's buffers stand for both)
        else:
            view._gode_full = full       # the solver passes this very object on to stage_fwd / vjp_phase1 / transform
        return view

    def _ybuf(self, y):
        """The [owned | halo] arena buffer behind a stage state handed out by ``new_Y`` (None for any other tensor)."""
        if not self.push_y or y is None:
            return None
        full = getattr(y, "_gode_full", None)
        if full is None and getattr(y, "_gode_y", False):
            full = y
        return full

    def _own_rows(self, full):
        return full[:self.plan.n_rows]

    def _transform_local(self, y, full, t, out):
        """S = transform(y) without an exchange of S: ``full`` holds this rank's rows of y and -- once the exchange issued by
        y's producer has completed -- the peers' rows in its halo tail."""
        n, nh = self.plan.n_rows, self.plan.halo.n_halo
        self.transform_rows(full, t, out, 0, n)          # owned rows: no wait
        self._wait(full)                                 # the peers' rows of y have landed (and: this reads the tail)
        if nh:
            self.transform_rows(full, t, out, n, nh)
        return out

    def numel_global(self):
        return self.plan.n_global * self.d

    def scalar(self, dev_scalar):
        if self.plan.world > 1:
            dist.all_reduce(dev_scalar, group=self.plan.group)
        v = float(dev_scalar.item())
        # the adaptive solver synchronises here once per step anyway: surface a timed-out peer wait on the hot path (a
        # time-out also traps the stream, peer.cu:k_peer_wait, so fixed-step solves fail at their next synchronisation)
        self.plan.check_peers()
        return v

    def reduce_small(self, t):
        if self.plan.world > 1:
            if t.is_contiguous():
                dist.all_reduce(t, group=self.plan.group)
            else:
                c = t.contiguous()
                dist.all_reduce(c, group=self.plan.group)
                t.copy_(c)
        return t

    # ---- halo exchange: peer-memory push, or NCCL (synchronous / on the side stream) ------------------------------
    def _exchange(self, halo, buf):
        plan = self.plan
        if self.peer is not None:
            self.pending[buf.data_ptr()] = self.peer.push(halo, buf)
            return
        if plan.mode == "sync" or plan.world == 1:
            halo.exchange(buf)
            return
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        cs = plan.comm_stream
        cs.wait_event(ready)
        buf.record_stream(cs)        # the allocator must not recycle the buffer under the side stream
        with torch.cuda.stream(cs):
            halo.exchange(buf)
            done = torch.cuda.Event()
            done.record(cs)
        self.pending[buf.data_ptr()] = done

    def _await(self, buf):
        """Enqueue the wait for ``buf``'s in-flight exchange (if any) without stamping the read."""
        ev = self.pending.pop(buf.data_ptr(), None)
        if self.peer is not None:
            if ev is not None:
                self.peer.wait(ev)       # ev is the exchange's epoch
            return
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def _wait(self, buf):
        """... and stamp it: a gather from ``buf`` is the NEXT thing enqueued on the consumer stream.  A kernel that gathers
        AND produces a fused push calls ``_await`` and passes the buffer as ``reads`` to ``_fused_begin`` instead, so that
        the stamp is taken after that call's hazard exchanges (peer.ExchangeProtocol.begin)."""
        self._await(buf)
        if self.peer is not None:
            self.peer.note_read(buf)     # a gather from buf is about to be enqueued

    def _own_pass(self, csr, X):
        """partial = (owned-column block) @ X[:n_own]  -- runs while X's halo tail is still arriving."""
        if self.partial is None:
            self.partial = self.new()
        nb = lib.gode_spmm_workspace_bytes(C.byref(csr), self.d)
        ws = ops.workspace(nb, self.dev, "spmm") if nb else None
        check(lib.gode_spmm_csr_f32(C.byref(csr), ops._p(X), self.d, self.d, ops._p(self.partial), self.d, None,
                                    ops._p(ws), nb, ops._stream()), "gode_spmm_csr_f32")
        return self.partial

    def _with_halo_pass(self, part, call):
        self.f_halo.partial_in = part.data_ptr()
        full, self.f = self.f, self.f_halo
        try:
            return call()
        finally:
            self.f = full

    def _push_fusable(self):
        return bool(lib.gode_gcn_push_fusable(C.byref(self.f)))

    def _fused_begin(self, field, halo, buf, also=(), reads=()):
        """The next producer of ``buf`` also stores the peers' rows (gode_push_route_t in the descriptor).  ``also``:
        (field, halo, buffer) of further operands the same kernel produces (one epoch for all); ``reads``: halo operands
        it gathers from."""
        epoch, _ = self.peer.begin(buf, also=[b for _, _, b in also], reads=reads)
        setattr(self.f, field, self.peer.fused_route(halo, buf))
        for f2, h2, b2 in also:
            setattr(self.f, f2, self.peer.fused_route(h2, b2))
        return epoch

    def _fused_end(self, field, buf, epoch, also=()):
        from . import _lib
        setattr(self.f, field, _lib.PushRoute())
        for f2, _, _ in also:
            setattr(self.f, f2, _lib.PushRoute())
        e = self.peer.finish(epoch)
        self.pending[buf.data_ptr()] = e
        for _, _, b2 in also:
            self.pending[b2.data_ptr()] = e

    def _transform_pipelined(self, y, t, out):
        """S = transform(y, t) in row chunks on the current stream, each chunk's boundary rows pushed on the side stream."""
        bounds = self.peer.part_bounds(self.pipe_S)
        ws = self._ws()

        def produce(c):
            check(lib.gode_gcn_transform_rows(C.byref(self.f), ops._p(y), float(t), ops._p(out), bounds[c],
                                              bounds[c + 1] - bounds[c], ops._p(ws), self.ws_bytes, ops._stream()),
                  "gode_gcn_transform_rows")

        self.pending[out.data_ptr()] = self.peer.push_pipelined(self.plan.halo, out, self.pipe_S, produce)
        return out

    def transform(self, y, t, out):
        full = self._ybuf(y)
        if full is not None and full.data_ptr() in self.pending:
            return self._transform_local(y, full, t, out)     # y's producer pushed the peers' rows (vjp_phase1)
        if self.pipe_S:
            return self._transform_pipelined(y, t, out)
        if self.fused_S:
            epoch = self._fused_begin("push_S", self.plan.halo, out)
            super().transform(y, t, out)
            self._fused_end("push_S", out, epoch)
            return out
        super().transform(y, t, out)
        self._exchange(self.plan.halo, out)
        return out

    def _stage_pipelined(self, S, k_out, y0, kprev, coefs, coef_self, y_next, t_next, S_next, second=None):
        """One forward stage in ``pipe_G`` row chunks: gather chunk c (+ its Runge-Kutta combination) and transform its rows
        on the main stream; push chunk c's boundary rows of S_next on the side stream underneath chunk c+1's gather."""
        from . import odeint as _od
        self.nfe += 1
        bounds = self.peer.part_bounds(self.pipe_G)
        ws = self._ws()
        if _od.MASK_LOG is not None and k_out is None:
            k_out = self.new()
        karr = (C.c_void_p * MAX_STAGES)(*[k.data_ptr() for k in kprev])
        carr = (C.c_float * MAX_STAGES)(*[float(c) for c in coefs])

        def produce(c):
            r0, nr = bounds[c], bounds[c + 1] - bounds[c]
            check(lib.gode_gcn_stage_fwd_rows(C.byref(self.f), ops._p(S), ops._p(k_out), ops._p(y0), karr, carr, len(kprev),
                                              float(coef_self), ops._p(y_next), r0, nr, ops._p(ws), self.ws_bytes,
                                              ops._stream()), "gode_gcn_stage_fwd_rows")
            check(lib.gode_gcn_transform_rows(C.byref(self.f), ops._p(y_next), float(t_next), ops._p(S_next), r0, nr,
                                              ops._p(ws), self.ws_bytes, ops._stream()), "gode_gcn_transform_rows")

        with self._second(second, len(kprev)):
            self.pending[S_next.data_ptr()] = self.peer.push_pipelined(self.plan.halo, S_next, self.pipe_G, produce, reads=[S])
        if _od.MASK_LOG is not None:
            _od.MASK_LOG.append(k_out > 0)

    def stage_fwd(self, S, k_out, y0=None, kprev=(), coefs=(), coef_self=0.0, y_next=None, t_next=0.0, S_next=None,
                  second=None):
        run = lambda: super(HaloKernelMixin, self).stage_fwd(S, k_out, y0, kprev, coefs, coef_self, y_next, t_next, S_next,
                                                             second=second)
        full = self._ybuf(y_next) if S_next is not None else None
        if full is not None:
            self._await(S)
            epoch = self._fused_begin("push_y", self.plan.halo, full, reads=[S])      # the gather's epilogue pushes y_next
            super(HaloKernelMixin, self).stage_fwd(S, k_out, y0, kprev, coefs, coef_self, y_next, t_next, None, second=second)
            self._fused_end("push_y", full, epoch)
            return self._transform_local(y_next, full, t_next, S_next)
        if self.pipe_G and S_next is not None and y_next is not None:
            self._await(S)
            return self._stage_pipelined(S, k_out, y0, kprev, coefs, coef_self, y_next, t_next, S_next, second)
        if self.pipe_S:
            self._wait(S)
            super(HaloKernelMixin, self).stage_fwd(S, k_out, y0, kprev, coefs, coef_self, y_next, t_next, None,
                                                   second=second)
            if S_next is not None:
                self._transform_pipelined(y_next, t_next, S_next)
            return
        if self.fused_S:
            if S_next is None:
                self._wait(S)
                return run()
            self._await(S)
            epoch = self._fused_begin("push_S", self.plan.halo, S_next, reads=[S])   # the transform inside stage_fwd pushes S_next
            run()
            return self._fused_end("push_S", S_next, epoch)
        if self.plan.split is None:
            self._wait(S)
            run()
        else:
            part = self._own_pass(self.csr_own, S)
            self._wait(S)
            self._with_halo_pass(part, run)
        if S_next is not None:
            self._exchange(self.plan.halo, S_next)

    def vjp_phase1(self, S, a, sign, k_y, gP, y0=None, kprev=(), coefs=(), coef_self=0.0, y_next=None, second=None):
        run = lambda: super(HaloKernelMixin, self).vjp_phase1(S, a, sign, k_y, gP, y0, kprev, coefs, coef_self, y_next,
                                                              second=second)
        if self.fused:
            self._await(S)
            full = self._ybuf(y_next)
            # the gather's epilogue pushes gP -- and y_next, when the next transform will be local -- under ONE epoch
            also = [("push_y", self.plan.halo, full)] if full is not None else []
            epoch = self._fused_begin("push_gP", self.plan.halo_t, gP, also=also, reads=[S])
            run()
            return self._fused_end("push_gP", gP, epoch, also=also)
        if self.plan.split is None:
            self._wait(S)
            run()
        else:
            part = self._own_pass(self.csr_own, S)
            self._wait(S)
            self._with_halo_pass(part, run)
        self._exchange(self.plan.halo_t, gP)

    def vjp_phase2(self, y, t, gP, k_a, gtheta, a0=None, kprev=(), coefs=(), coef_self=0.0, a_next=None, second=None):
        run = lambda: super(HaloKernelMixin, self).vjp_phase2(y, t, gP, k_a, gtheta, a0, kprev, coefs, coef_self, a_next,
                                                              second=second)
        if self.plan.split is None:
            self._wait(gP)
            return run()
        part = self._own_pass(self.csr_own_t, gP)
        self._wait(gP)
        return self._with_halo_pass(part, run)


class PartitionedPlan:
    """This rank's row block of A_hat and of A_hat^T with their halo plans (what ``GcnKernel`` needs from a plan).

    ``mode`` (argument, else ``GODE_HALO_MODE``, else "async" when world > 1):
      "sync"   every exchange is issued and awaited on the compute stream;
      "async"  exchanges run on a high-priority side stream and are awaited by the kernel that gathers from the
               buffer -- the adjoint issues the next stage's transform before phase 2 (odeint._gcn_aug_fixed_step), so
               both of its exchanges hide under the transform and the dense VJP chain; the 4 forward exchanges of an
               rk4 step have no independent work to hide under and stay exposed;
      "p2p"    no NCCL and no send buffer on the data path: one libgode kernel stores the rows a peer references straight
               into that peer's halo tail over NVLink peer mappings and publishes an epoch flag; the gather waits for the
               peers' flags on the device (peer.py, csrc/peer.cu).  Operand buffers come from a cudaIpc-shared arena;
      "p2p-async"  the same with the push on the high-priority side stream, so it runs underneath independent kernels
               exactly as "async" does for the NCCL exchange;
      "p2p-fused"  no exchange kernel at all: the kernels that PRODUCE a gather operand (the tensor-core transform for S,
               the phase-1 gather's epilogue for gP) store every row a peer references into that peer's halo tail next
               to their own local store (posted NVLink writes riding underneath the kernel), then a one-CTA kernel
               publishes the epoch flag.  Falls back to "p2p" for shapes the tensor-core transform does not cover;
      "split"  "async" plus a column split of each block -- ``A_own`` [n_own, n_own] and ``A_halo`` [n_own, n_halo] --
               so a gather runs in two passes: owned columns while the halo is in flight, then halo columns with the
               first pass as ``partial_in``.  Measured at 2 GPUs (N = 10 M): the second pass over all rows costs more
               (+33 ms/step) than the exchange it hides; kept for graphs with a small boundary.
    """

    def __init__(self, n, bounds, rank, A, At, halo, halo_t, nnz_global, group=None, split=None, mode="sync"):
        self.n_global, self.bounds, self.rank, self.group = n, bounds, rank, group
        self.world = len(bounds) - 1
        self.lo, self.hi = bounds[rank], bounds[rank + 1]
        self.n_rows = self.hi - self.lo
        self.A, self.At, self.halo, self.halo_t = A, At, halo, halo_t
        self.nnz = A.nnz
        self.nnz_global = nnz_global
        self.rowptr_t = At.rowptr            # "has a transpose" marker GcnKernel looks at
        self.device = A.device
        self.split = split                   # None, or dict(A_own, A_halo, At_own, At_halo)
        self.mode = mode
        self.comm_stream = (torch.cuda.Stream(device=self.device, priority=-1)
                            if (mode not in ("sync", "p2p") and A.device.type == "cuda") else None)
        self._peer = {}                      # feature width -> peer.PeerHalo (modes "p2p", "p2p-async")
        if self.comm_stream is not None:
            # the exchange of gP is issued right before the (persistent, one CTA per SM) transform kernel: without free
            # SMs the NCCL kernel only starts when that kernel ends (torch.profiler timeline, profiles/r01_halo_overlap.md).
            # Measured at 2 GPUs: reserving 20 SMs raises the overlapped NCCL time from 6.2 to 10.6 ms per step but slows
            # the persistent kernels and NCCL itself by more (162.8 vs 159.2 ms per step) -> default 0, knob kept.
            import os
            check(lib.gode_reserve_sms(int(os.environ.get("GODE_RESERVE_SMS", "0"))), "gode_reserve_sms")

    @classmethod
    def build(cls, row, col, val, n, rank=None, world=None, group=None, bounds=None, mode=None):
        """``row, col, val``: the COO of A_hat (int64, int64, fp32) -- at least every entry whose row or column
        this rank owns; the full COO is fine."""
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        import os
        # default on GPUs: the producers push boundary rows into the peers' halo tails themselves ("p2p-fused"; it falls
        # back to "p2p" for shapes without a tensor-core transform and to "async" when peer mappings are unavailable)
        on_gpu = row.device.type == "cuda"
        mode = mode or os.environ.get("GODE_HALO_MODE") or (("p2p-fused" if on_gpu else "async") if world > 1 else "sync")
        if mode not in MODES:
            raise ValueError("halo mode must be one of %s" % (MODES,))
        if world == 1:
            mode = "sync"
        overlap = mode == "split"
        bounds = bounds or partition_bounds(n, world)
        n_own = bounds[rank + 1] - bounds[rank]
        split = {} if overlap else None
        r, c, v, halo = local_block(row, col, val, bounds, rank)
        A = ops.GraphPlan.from_coo(r, c, v, n_own, n_own + int(halo.numel()), build_transpose=False)
        if overlap:
            own, hal = split_block(r, c, v, n_own)
            split["A_own"] = ops.GraphPlan.from_coo(*own, n_own, n_own, build_transpose=False)
            split["A_halo"] = ops.GraphPlan.from_coo(*hal, n_own, max(int(halo.numel()), 1), build_transpose=False)
            del own, hal
        del r, c, v
        r, c, v, halo_t = local_block(row, col, val, bounds, rank, transpose=True)
        At = ops.GraphPlan.from_coo(r, c, v, n_own, n_own + int(halo_t.numel()), build_transpose=False)
        if overlap:
            own, hal = split_block(r, c, v, n_own)
            split["At_own"] = ops.GraphPlan.from_coo(*own, n_own, n_own, build_transpose=False)
            split["At_halo"] = ops.GraphPlan.from_coo(*hal, n_own, max(int(halo_t.numel()), 1), build_transpose=False)
            del own, hal
        del r, c, v
        return cls(n, bounds, rank, A, At, HaloPlan(halo, bounds, rank, group), HaloPlan(halo_t, bounds, rank, group),
                   int(val.numel()), group, split, mode)

    def csr(self, transpose=False):
        return (self.At if transpose else self.A).csr(False)

    def unit_transpose(self):
        """(row scale of the owned rows, 0/1 pattern of this rank's A_hat^T block) when EVERY rank's block of A_hat is
        row-constant and row-stochastic (``ops.GraphPlan.unit_transpose``: the owner of a row pre-scales its gP row, so
        the halo rows of gP arrive scaled and all ranks have to agree) and gathers are not column-split; else None.
        Collective on first use."""
        if not hasattr(self, "_unit_t"):
            scale = None if self.split is not None else self.A.row_stochastic_scale()
            ok = torch.tensor([1 if scale is not None else 0], dtype=torch.int32, device=self.device)
            if self.world > 1 and dist.is_initialized():
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            self._unit_t = (scale, self.At.unit_pattern_csr()) if int(ok.item()) == 1 else None
        return self._unit_t

    def peer_for(self, d):
        """The peer-memory arena / exchange state for feature width ``d`` (collective on first use)."""
        if not self.mode.startswith("p2p") or self.world == 1 or self.device.type != "cuda":
            return None
        ph = self._peer.get(d)
        if ph is None:
            from . import peer
            try:
                import os
                side = self.mode == "p2p-async" or (self.mode == "p2p-fused" and (
                    os.environ.get("GODE_PIPE_S", "0") != "0" or pipe_g_chunks(self.world) > 0))
                ph = self._peer[d] = peer.PeerHalo(self, d, push_stream=self.comm_stream if side else None)
            except peer.PeerSetupError as e:
                # raised on every rank together: all of them switch to the NCCL exchange on the side stream
                import warnings
                warnings.warn("%s -- falling back to halo mode 'async' (NCCL)" % e)
                self.mode = "async"
                if self.comm_stream is None:
                    self.comm_stream = torch.cuda.Stream(device=self.device, priority=-1)
                return None
        return ph

    def check_peers(self):
        """Raises if a device-side wait of the peer-memory exchange timed out (synchronises the stream)."""
        for ph in self._peer.values():
            ph.check()

    def make_kernel(self, *args):
        from .odeint import GcnKernel
        cls = type("PartitionedGcnKernel", (HaloKernelMixin, GcnKernel), {})
        return cls(self, *args)

    def halo_bytes_per_step(self, d, n_fwd_evals, n_aug_evals):
        """Bytes this rank moves over NVLink (sent + received) in one fixed-step fwd+bwd step: one support exchange per
        forward evaluation and per augmented evaluation, one gP exchange per augmented evaluation (``n_aug_evals`` counts
        the VJP evaluations; the evaluation of f(t1) that opens the adjoint shares its support with the first of them)."""
        s = self.halo.bytes_per_exchange(d)
        g = self.halo_t.bytes_per_exchange(d)
        return (n_fwd_evals + n_aug_evals) * s + n_aug_evals * g


# --------------------------------------------------------------------------------------------------
# data-parallel molecule batches (BASELINE config 5): independent units, one gradient all-reduce per step
# --------------------------------------------------------------------------------------------------


def shard_molecules(n_mol, rank, world):
    """Contiguous shard [lo, hi) of the molecules of a batch for this rank (sizes differ by at most one)."""
    b = partition_bounds(n_mol, world)
    return b[rank], b[rank + 1]


def allreduce_gradients(params, group=None, local_weight=1.0):
    """Sum ``local_weight * grad`` over ranks, in place, through one flat buffer (one NCCL all-reduce per step; the QC
    model has 14.3 M parameters = 57 MB).  With a batch-mean loss pass ``local_weight = local_batch / global_batch``
    so the result is the gradient of the global-batch mean (QC/util.py:184 uses MSELoss's mean).

    The buffer covers EVERY parameter that requires a gradient, with zeros for those whose ``grad`` is None on this rank
    (an unused branch, an empty shard): all ranks therefore always issue the same-sized collective, whatever their local
    graph looked like (a rank that skipped the call, or sent a shorter buffer, would hang or corrupt the all-reduce)."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    if local_weight != 1.0:
        flat.mul_(local_weight)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, group=group)
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            p.grad = flat[off:off + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def allreduce_mean(total, count, device, group=None):
    """Global mean of a per-rank (sum, count) pair -- validation metrics of data-parallel runs (every rank then takes the
    same checkpoint / early-stopping decision)."""
    t = torch.tensor([float(total), float(count)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)
    return float(t[0] / t[1].clamp(min=1.0))


# --------------------------------------------------------------------------------------------------
# locality-aware relabelling ("degree/locality-aware reordering, a pure relabelling")
# --------------------------------------------------------------------------------------------------


def _csr_pattern(row, col, n):
    """rowptr / colidx of the (row-sorted) pattern; ``row`` need not be sorted."""
    order = torch.argsort(row, stable=True)
    colidx = col[order]
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=row.device)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=n), 0)
    return rowptr, colidx


def locality_order(row, col, n, method="rcm", sweeps=0):
    """A relabelling of the nodes of a symmetric pattern that shortens the distance |i - j| of connected nodes, so that a
    contiguous row partition cuts fewer edges (smaller halo) and a row tile's neighbours share L2 lines.

    Returns ``perm`` (int64 [n]): the NEW id of old node ``i`` is ``perm[i]``.  A relabelling is parity-neutral: the
    relabelled problem is the same problem (``relabel`` / ``unrelabel`` are exact inverses -- tests/test_parallel_cpu.py).

    ``method``:
      * ``"degree"`` -- nodes by descending degree (hubs first: the rows everybody references share one partition and
        stay L2-resident);
      * ``"rcm"``    -- reverse Cuthill-McKee: level-synchronous breadth-first search on the device from a minimum-degree
        node of every component; inside a level nodes are ordered by the position of their first-discovered parent, then
        by degree; the final order is reversed.
    ``sweeps`` > 0 adds barycentre sweeps: every node moves to the mean position of its neighbours and the nodes are
    re-ranked -- the classical one-dimensional placement refinement; it tightens a band that the breadth-first levels
    found coarsely.
    """
    dev = row.device
    row, col = row.to(torch.int64), col.to(torch.int64)
    deg = torch.bincount(row, minlength=n)
    if method == "degree":
        order = torch.argsort(deg, descending=True, stable=True)          # order[k] = old id at new position k
    elif method == "rcm":
        rowptr, colidx = _csr_pattern(row, col, n)
        pos = torch.full((n,), -1, dtype=torch.int64, device=dev)          # CM position of each node
        placed = 0
        # components in order of their minimum-degree node
        start_order = torch.argsort(deg, stable=True)
        sp = 0
        while placed < n:
            while sp < n and int(pos[start_order[sp]]) >= 0:
                sp += 1
            frontier = start_order[sp:sp + 1]
            pos[frontier] = placed
            placed += 1
            while frontier.numel():
                cnt = rowptr[frontier + 1] - rowptr[frontier]
                tot = int(cnt.sum())
                if tot == 0:
                    break
                # expand: (parent position, child) for every stored entry of the frontier rows
                owner = torch.repeat_interleave(torch.arange(frontier.numel(), device=dev), cnt)
                offs = torch.arange(tot, device=dev) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt)
                child = colidx[rowptr[frontier][owner] + offs]
                ppos = pos[frontier][owner]
                new = pos[child] < 0
                child, ppos = child[new], ppos[new]
                if child.numel() == 0:
                    break
                # first-discovered parent of every new node, then order by (parent position, degree, id)
                uniq, inv = torch.unique(child, return_inverse=True)
                first = torch.full((uniq.numel(),), 1 << 62, dtype=torch.int64, device=dev)
                first.scatter_reduce_(0, inv, ppos, reduce="amin")
                key = torch.argsort(deg[uniq], stable=True)
                key = key[torch.argsort(first[key], stable=True)]
                frontier = uniq[key]
                pos[frontier] = placed + torch.arange(frontier.numel(), device=dev)
                placed += int(frontier.numel())
        order = torch.argsort(pos, stable=True).flip(0)                    # reverse Cuthill-McKee
    else:
        raise ValueError("unknown reordering method %r" % (method,))
    perm = torch.empty(n, dtype=torch.int64, device=dev)
    perm[order] = torch.arange(n, device=dev)
    for _ in range(int(sweeps)):
        p = perm.to(torch.float64)
        s = torch.zeros(n, dtype=torch.float64, device=dev).index_add_(0, row, p[col])
        bary = torch.where(deg > 0, s / deg.clamp(min=1), p)
        order = torch.argsort(bary, stable=True)
        perm[order] = torch.arange(n, device=dev)
    return perm


def relabel(row, col, perm):
    """The same entries under the new ids (values travel with their entries unchanged)."""
    return perm[row], perm[col]


def unrelabel(row, col, perm):
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(perm.numel(), device=perm.device)
    return inv[row], inv[col]


def halo_fraction(row, col, n, world):
    """Halo rows per owned row, averaged over the ranks of a contiguous ``world``-way row partition (the quantity the
    relabelling is meant to shrink)."""
    b = torch.tensor(partition_bounds(n, world), device=row.device)
    ro, co = torch.bucketize(row, b[1:-1], right=True), torch.bucketize(col, b[1:-1], right=True)
    cut = ro != co
    pairs = torch.unique(ro[cut] * n + col[cut])
    return float(pairs.numel()) / n
