"""Halo exchange over NVLink peer memory (csrc/peer.cu) for the row-partitioned GCN-ODE path.

The reference is single-device; parallel.py partitions its ``torch.spmm(self.adj, support)``
(GCN/layers.py:71) by rows.  This module replaces "pack rows -> NCCL all-to-all-v -> halo tail" by one libgode kernel
that stores the rows a peer references straight into the halo tail of that peer's operand buffer (``gode_halo_push``)
and a device-side wait on the peers' epoch flags (``gode_peer_wait``) in front of the kernel that gathers.

Operand buffers (supports S, masked adjoints gP) therefore live in one ``cudaMalloc`` arena per rank that every peer
maps once (cudaIpc handles shipped with ``all_gather_object``); slots have the same offsets on every rank.

Write-after-read across ranks.  A push overwrites the halo tail of the peers' slot X; the peers may still be gathering
from the previous contents of X.  All ranks run the same program, so the rule is decided locally and identically
everywhere: a gather from X is stamped with the number of exchanges issued when the gather is ENQUEUED (``c``: for a
kernel that gathers and produces a fused push, after that call's own hazard exchanges -- ``ExchangeProtocol.begin``);
a later push into X needs a completed wait on an exchange ``j >= c + 1`` to have been enqueued first (a peer that has
published ``j`` has finished everything it enqueued before its push ``j``, in particular that gather).  If none was, one is enqueued (an empty
exchange when no exchange was issued since the gather).  The solver's buffer rotation makes this the rare path.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import ops
from ._lib import lib, check, PeerGroup, PushRoute, MAX_PEERS, PEER_HEADER_BYTES, PEER_HANDLE_BYTES


class PeerSetupError(RuntimeError):
    """Raised on EVERY rank when the peer-memory arena could not be set up on some rank."""


class SlotPool:
    """Equally sized arena slots handed out round-robin (the free slot after the last one handed out): deterministic
    across ranks, and a slot is reused as late as possible, which keeps pushes clear of the peers' last gathers."""

    def __init__(self, n_slots):
        self.free = set(range(n_slots))
        self.n_slots = n_slots
        self.cursor = 0

    def acquire(self):
        if not self.free:
            raise RuntimeError("peer arena exhausted: more than %d halo operand buffers alive (raise GODE_PEER_SLOTS)"
                               % self.n_slots)
        for k in range(self.n_slots):
            slot = (self.cursor + k) % self.n_slots
            if slot in self.free:
                self.free.discard(slot)
                self.cursor = (slot + 1) % self.n_slots
                return slot

    def release(self, slot):
        self.free.add(slot)


class HazardTracker:
    """The cross-rank write-after-read rule of the module docstring as pure bookkeeping (unit-tested on the CPU)."""

    def __init__(self):
        self.issued = 0        # exchanges issued so far (= the epoch of the latest push)
        self.waited = 0        # highest epoch whose wait has been enqueued on the consumer stream
        self.read_at = {}      # slot -> value of ``issued`` when the slot was last gathered from

    def note_read(self, slot):
        self.read_at[slot] = self.issued

    def before_push(self, slot):
        """Returns (n_empty_exchanges, wait_epoch or None) to enqueue before pushing into ``slot``."""
        c = self.read_at.pop(slot, None)
        if c is None or self.waited >= c + 1:
            return 0, None
        n_empty = 0
        if self.issued < c + 1:
            n_empty = 1
            self.issued += 1
        self.waited = self.issued
        return n_empty, self.issued

    def push(self):
        self.issued += 1
        return self.issued

    def wait(self, epoch):
        """Returns True when a wait kernel for ``epoch`` still has to be enqueued."""
        if self.waited >= epoch:
            return False
        self.waited = epoch
        return True


class ExchangeProtocol:
    """Order of pushes, waits and cross-stream events of the peer-memory exchange, over abstract primitives (the CUDA
    ones in PeerHalo; tests/test_peer_protocol_cpu.py replays them in a randomised multi-rank simulator).

    Every push (also an empty one) goes to ONE stream -- the side stream when ``side`` is set, else the consumer
    stream -- so the epoch flags a peer sees never decrease.  Hazard waits go to the consumer stream and the push is
    ordered after them and after the producer of the buffer."""

    side = False

    def __init__(self, n_slots):
        self.pool = SlotPool(n_slots)
        self.track = HazardTracker()
        self.done = set()                      # epochs whose own push (side stream) no consumer has waited for yet
        self.own_side = {}                     # slot -> epoch of this rank's last side-stream push into it (while in ``done``)

    # primitives ---------------------------------------------------------------------------------------------------
    def _slot(self, buf):
        raise NotImplementedError

    def _emit_push(self, epoch, halo, buf, slot):
        raise NotImplementedError

    def _emit_wait(self, epoch):
        raise NotImplementedError

    def _emit_push_part(self, epoch, halo, buf, slot, part, n_parts):
        raise NotImplementedError

    def _emit_side_after_main(self):
        """side stream waits for everything enqueued on the consumer stream so far"""

    def _emit_done(self, epoch):
        """record the completion of this rank's push ``epoch`` on the side stream"""

    def _emit_wait_done(self, epoch):
        """consumer stream waits for the recorded completion of push ``epoch``"""

    # protocol -----------------------------------------------------------------------------------------------------
    def begin(self, buf, also=(), reads=()):
        """Hazard handling for an exchange into ``buf``'s slot (and the slots of ``also``: further buffers the SAME producer
        kernel fills, sharing the epoch); returns (epoch, slot).  What follows is either ``_emit_push`` of the data (``push``)
        or a producer kernel that stores the peers' rows itself, then ``finish``.

        ``reads``: halo operands that producer kernel itself GATHERS from.  They are stamped here -- after the hazard
        exchanges of this call, before the epoch is assigned -- because the stamp has to name the last epoch published
        BEFORE the gather is enqueued: an empty exchange inserted by the hazard rule is published ahead of the gather, so a
        peer that has published it has not necessarily finished that gather (found by replaying the adaptive solver, whose
        single gP buffer makes the rule act, through tests/test_peer_protocol_cpu.py; stamping at the wait in front of the
        call let a faster rank overwrite the support a slower rank was still gathering from)."""
        for b in (buf,) + tuple(also):
            # write-after-write inside this rank: an earlier push of OURS into the same slot may still be running on the side
            # stream (the buffer was re-produced, or released and its slot re-allocated, without a wait in between); a fused
            # producer stores from the consumer stream, so that stream first waits for the old push (found by
            # tests/test_peer_protocol_cpu.py::test_random_api_programs_are_safe)
            e_own = self.own_side.pop(self._slot(b), None)
            if e_own is not None and e_own in self.done:
                self._emit_wait_done(e_own)
                self.done.discard(e_own)
            n_empty, wait_epoch = self.track.before_push(self._slot(b))
            if n_empty:
                if self.side:
                    self._emit_side_after_main()
                self._emit_push(self.track.issued, None, None, None)
            if wait_epoch is not None:
                self._emit_wait(wait_epoch)
        for r in reads:
            self.track.note_read(self._slot(r))
        return self.track.push(), self._slot(buf)

    def push(self, halo, buf):
        """Issue the exchange that fills the peers' halo tails of ``buf``'s slot; returns its epoch."""
        epoch, slot = self.begin(buf)
        if self.side:
            self._emit_side_after_main()
        self._emit_push(epoch, halo, buf, slot)
        if self.side:
            self._emit_done(epoch)
            self.done.add(epoch)
            self.own_side[slot] = epoch
        return epoch

    def finish(self, epoch):
        """Fused exchange: the producer kernel (consumer stream) has stored the peers' rows; publish the epoch -- on the
        push stream, like every other flag write, ordered after the producer."""
        if self.side:
            self._emit_side_after_main()
        self._emit_push(epoch, None, None, None)
        if self.side:
            self._emit_done(epoch)
            self.done.add(epoch)
        return epoch

    def push_pipelined(self, halo, buf, n_parts, produce, reads=()):
        """One exchange in ``n_parts`` row chunks: ``produce(c)`` enqueues the producer of chunk c on the consumer stream,
        the rows of that chunk are pushed on the side stream while the next chunk is produced; the last part publishes
        the epoch.  Needs the side stream (flag writes stay on one stream)."""
        assert self.side, "the pipelined exchange needs the side stream"
        epoch, slot = self.begin(buf, reads=reads)      # ``produce`` may gather from ``reads`` (see begin)
        for c in range(n_parts):
            produce(c)
            self._emit_side_after_main()
            self._emit_push_part(epoch, halo, buf, slot, c, n_parts)
        self._emit_done(epoch)
        self.done.add(epoch)
        self.own_side[slot] = epoch
        return epoch

    def wait(self, epoch):
        """Enqueue (once) the wait for exchange ``epoch`` on the consumer stream: the peers' flags, and -- with a side
        stream -- this rank's own push (it reads the buffer the next producer will overwrite)."""
        for e in sorted(e for e in self.done if e <= epoch):
            self._emit_wait_done(e)
            self.done.discard(e)
        if self.track.wait(epoch):
            self._emit_wait(epoch)

    def note_read(self, buf):
        """A gather from ``buf`` is about to be enqueued on the consumer stream."""
        self.track.note_read(self._slot(buf))


class _Slot:
    """Arena memory exposed through __cuda_array_interface__; the slot returns to the pool when the last tensor
    (or view) over it dies."""

    def __init__(self, owner, slot, ptr, rows, d):
        self.owner, self.slot = owner, slot
        self.__cuda_array_interface__ = {"shape": (rows, d), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                         "strides": None}

    def __del__(self):
        try:
            self.owner._release(self.slot)
        except Exception:
            pass


class PeerHalo(ExchangeProtocol):
    """Arena + peer mappings + exchange bookkeeping for one (partitioned plan, feature width)."""

    def __init__(self, plan, d, n_slots=None, timeout_s=None, max_ctas=None, push_stream=None):
        if plan.device.type != "cuda":
            raise TypeError("peer-memory halo exchange needs CUDA devices")
        if plan.world > MAX_PEERS:
            raise ValueError("peer-memory halo exchange supports at most %d ranks" % MAX_PEERS)
        self.plan, self.d = plan, d
        self.world, self.rank, self.group = plan.world, plan.rank, plan.group
        dev = plan.device
        n_slots = n_slots or int(os.environ.get("GODE_PEER_SLOTS", "12"))    # S x3, gP x2, stage states x2 (+ multi-step grids)
        self.timeout_ns = int(float(timeout_s or os.environ.get("GODE_PEER_TIMEOUT_S", "30")) * 1e9)
        self.max_ctas = int(max_ctas or os.environ.get("GODE_PUSH_CTAS", "0"))
        # slot size: the largest operand buffer over all ranks (same offsets everywhere)
        rows = torch.tensor([plan.n_rows + max(plan.halo.n_halo, plan.halo_t.n_halo)], dtype=torch.int64, device=dev)
        dist.all_reduce(rows, op=dist.ReduceOp.MAX, group=self.group)
        self.slot_rows = int(rows.item())
        self.slot_bytes = -(-self.slot_rows * d * 4 // 4096) * 4096
        self.n_slots = n_slots
        nbytes = PEER_HEADER_BYTES + n_slots * self.slot_bytes
        # Every step that can fail locally (allocation, export, mapping a peer's arena) is followed by an agreement over
        # all ranks, so that either every rank ends up with a working exchange or every rank raises PeerSetupError and
        # the caller falls back to the NCCL exchange -- never a mix that would hang in a collective.
        self.base, self.opened = None, []
        base = C.c_void_p()
        handle = (C.c_ubyte * PEER_HANDLE_BYTES)()
        err = None
        try:
            check(lib.gode_peer_alloc(nbytes, C.byref(base)), "gode_peer_alloc")
            self.base = base.value
            check(lib.gode_peer_export(self.base, handle), "gode_peer_export")
        except Exception as e:          # noqa: BLE001 -- reported through PeerSetupError on every rank
            err = e
        self._agree(err, "allocating / exporting the arena")
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        g = PeerGroup()
        g.world, g.rank = self.world, self.rank
        try:
            for p in range(self.world):
                if p == self.rank:
                    g.base[p] = self.base
                    continue
                h = (C.c_ubyte * PEER_HANDLE_BYTES).from_buffer_copy(handles[p])
                out = C.c_void_p()
                check(lib.gode_peer_open(h, C.byref(out)), "gode_peer_open")
                g.base[p] = out.value
                self.opened.append(out.value)
        except Exception as e:          # noqa: BLE001
            err = e
        self._agree(err, "mapping the peers' arenas")
        self.g = g
        ExchangeProtocol.__init__(self, n_slots)
        self.slot_of = {}                      # data_ptr -> slot
        self.push_stream = push_stream         # None: pushes on the current stream
        self.side = push_stream is not None
        self.done_events = {}                  # epoch -> event of this rank's own push on the side stream
        self._fused = {}                       # id(halo plan) -> (ptr, ent) device arrays of the fused push route
        self._part_cache = {}                  # (id(halo plan), n_parts) -> per-chunk segment ranges
        # where my rows land in every peer's halo tail, for both halo plans
        self.routes = {id(plan.halo): self._route(plan.halo), id(plan.halo_t): self._route(plan.halo_t)}
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)         # every arena is mapped and zeroed before the first flag is written

    def _agree(self, err, what):
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=self.plan.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            self.close()
            raise PeerSetupError("peer-memory halo exchange unavailable (%s failed on %s): %s" % (
                what, "this rank" if err is not None else "another rank", err))

    def _route(self, halo):
        """(send_ptr[world+1], dst_row[world]) host arrays: my rows for peer p go to rows dst_row[p] + k of its buffer."""
        dev = self.plan.device
        mine = torch.tensor([halo.n_own] + list(halo.recv_counts), dtype=torch.int64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allr, mine, group=self.group)
        allr = [t.tolist() for t in allr]
        send_ptr = [0]
        for p in range(self.world):
            send_ptr.append(send_ptr[-1] + int(halo.send_counts[p]))
        dst_row = []
        for p in range(self.world):
            n_own_p, recv_p = allr[p][0], allr[p][1:]
            assert recv_p[self.rank] == halo.send_counts[p], "halo plans disagree between ranks"
            dst_row.append(n_own_p + sum(recv_p[:self.rank]))
        return ((C.c_int64 * (self.world + 1))(*send_ptr), (C.c_int64 * self.world)(*dst_row))

    def fused_route(self, halo, buf):
        """_lib.PushRoute for the kernel that produces ``buf``: per owned row, the (peer, destination row) pairs."""
        key = id(halo)
        if key not in self._fused:
            send_ptr, dst_row = self.routes[key]
            dev = self.plan.device
            n_send = int(halo.send_idx.numel())
            sp = torch.tensor(list(send_ptr), dtype=torch.int64, device=dev)
            counts = sp[1:] - sp[:-1]
            peer_of = torch.repeat_interleave(torch.arange(self.world, device=dev, dtype=torch.int64), counts)
            k = torch.arange(n_send, device=dev, dtype=torch.int64) - sp[:-1][peer_of]
            dst = torch.tensor(list(dst_row), dtype=torch.int64, device=dev)[peer_of] + k
            ent = (peer_of << 40) | dst
            rows = halo.send_idx.to(torch.int64)
            order = torch.argsort(rows, stable=True)
            ptr = torch.zeros(halo.n_own + 1, dtype=torch.int64, device=dev)
            ptr[1:] = torch.cumsum(torch.bincount(rows, minlength=halo.n_own), 0)
            self._fused[key] = (ptr.to(torch.int32).contiguous(), ent[order].contiguous())
        ptr, ent = self._fused[key]
        r = PushRoute()
        r.ptr, r.ent = ptr.data_ptr(), (ent.data_ptr() if ent.numel() else ptr.data_ptr())
        off = PEER_HEADER_BYTES + self._slot(buf) * self.slot_bytes
        for p in range(self.world):
            if p != self.rank:
                r.base[p] = self.g.base[p] + off
        return r

    # ---- buffers ------------------------------------------------------------------------------------------------
    def new(self, rows):
        if rows > self.slot_rows:
            raise ValueError("operand buffer larger than an arena slot")
        slot = self.pool.acquire()
        ptr = self.base + PEER_HEADER_BYTES + slot * self.slot_bytes
        t = torch.as_tensor(_Slot(self, slot, ptr, max(rows, 1), self.d), device=self.plan.device)[:rows]
        assert t.data_ptr() == ptr or rows == 0
        self.slot_of[ptr] = slot
        return t

    def _release(self, slot):
        self.pool.release(slot)

    def _slot(self, buf):
        slot = self.slot_of.get(buf.data_ptr())
        if slot is None:
            raise ValueError("halo operand was not allocated from the peer arena (use kern.new_S() / kern.new_gP())")
        return slot

    # ---- exchange primitives (CUDA) -----------------------------------------------------------------------------
    def _push_stream_handle(self):
        return ops._stream() if self.push_stream is None else C.c_void_p(self.push_stream.cuda_stream)

    def _emit_push(self, epoch, halo, buf, slot):
        stream = self._push_stream_handle()
        if halo is None:
            zero = (C.c_int64 * (self.world + 1))()
            check(lib.gode_halo_push(C.byref(self.g), epoch, None, zero, zero, PEER_HEADER_BYTES, self.d, None, self.d,
                                     self.d, 1, stream), "gode_halo_push")
            return
        send_ptr, dst_row = self.routes[id(halo)]
        off = PEER_HEADER_BYTES + slot * self.slot_bytes
        check(lib.gode_halo_push(C.byref(self.g), epoch, ops._p(halo.send_idx), send_ptr, dst_row, off, self.d,
                                 ops._p(buf), buf.stride(0), self.d, self.max_ctas, stream), "gode_halo_push")

    def part_bounds(self, n_parts):
        """Row bounds of the chunks of a pipelined exchange (multiples of 128 rows: the transform's tile)."""
        n = self.plan.n_rows
        tiles = -(-n // 128)
        return [min(n, 128 * ((tiles * c) // n_parts)) for c in range(n_parts)] + [n]

    def _parts(self, halo, n_parts):
        """Per chunk: (seg_begin[world], seg_end[world], dst_row[world]) host arrays -- the entries of every peer's send
        segment (sorted by row) whose row falls into the chunk."""
        key = (id(halo), n_parts)
        if key not in self._part_cache:
            send_ptr, dst_row = self.routes[id(halo)]
            sp = list(send_ptr)
            rows = halo.send_idx.to(torch.int64)
            bounds = torch.tensor(self.part_bounds(n_parts), dtype=torch.int64, device=rows.device)
            cut = []                                    # cut[p][c]: first entry of peer p's segment with row >= bounds[c]
            for p_ in range(self.world):
                seg = rows[sp[p_]:sp[p_ + 1]]
                cut.append((torch.searchsorted(seg, bounds) + sp[p_]).tolist() if seg.numel() else [sp[p_]] * (n_parts + 1))
            out = []
            for c in range(n_parts):
                b = (C.c_int64 * self.world)(*[cut[p_][c] for p_ in range(self.world)])
                e = (C.c_int64 * self.world)(*[cut[p_][c + 1] for p_ in range(self.world)])
                dr = (C.c_int64 * self.world)(*[dst_row[p_] + cut[p_][c] - sp[p_] for p_ in range(self.world)])
                out.append((b, e, dr))
            self._part_cache[key] = out
        return self._part_cache[key]

    def _emit_push_part(self, epoch, halo, buf, slot, part, n_parts):
        b, e, dr = self._parts(halo, n_parts)[part]
        off = PEER_HEADER_BYTES + slot * self.slot_bytes
        check(lib.gode_halo_push_part(C.byref(self.g), epoch, ops._p(halo.send_idx), b, e, dr, off, self.d, ops._p(buf),
                                      buf.stride(0), self.d, self.max_ctas, 1 if part == n_parts - 1 else 0,
                                      self._push_stream_handle()), "gode_halo_push_part")

    def _emit_wait(self, epoch):
        check(lib.gode_peer_wait(C.byref(self.g), epoch, self.timeout_ns, ops._stream()), "gode_peer_wait")

    def _emit_side_after_main(self):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.push_stream.wait_event(ev)

    def _emit_done(self, epoch):
        ev = torch.cuda.Event()
        ev.record(self.push_stream)
        self.done_events[epoch] = ev

    def _emit_wait_done(self, epoch):
        torch.cuda.current_stream().wait_event(self.done_events.pop(epoch))

    def status(self):
        st = C.c_int32()
        check(lib.gode_peer_status(C.byref(self.g), C.byref(st), ops._stream()), "gode_peer_status")
        return st.value

    def check(self):
        if self.status() != 0:
            raise RuntimeError("peer halo exchange timed out waiting for a peer's flag (rank %d)" % self.rank)

    def close(self):
        for p in self.opened:
            lib.gode_peer_close(p)
        self.opened = []
        if self.base:
            lib.gode_peer_free(self.base)
            self.base = None
