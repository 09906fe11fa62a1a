"""Batch collate of molecule graphs (QC/datasets/utils.py:153-217), as the reference's ``collate_fn`` and on the device.

A molecule is the reference's dataset item ``((M, x, e), o)``: ``M`` the [n, n] adjacency matrix, ``x`` the n node
feature rows, ``e`` a dict ``{(src, tgt): feature row}`` with one entry per undirected edge, ``o`` the target row.

* ``collate_g_concat_edge_data(batch)`` -- the reference function (same argument, same 8-tuple
  ``(batch_size, G, B, X, E_d, E_src, E_tgt, Y)`` with the same dtypes), vectorised over numpy instead of a Python loop
  per edge.  ``dense=False`` returns ``E_tgt`` as the target-node index vector [2M] (what the layers here aggregate by;
  the dense one-hot [N, 2M] of the reference is 43 GB at 4096 molecules) and ``G`` as None.
* ``MoleculeStore`` -- the dataset as a ragged store on the device, built once; ``store.collate(ids)`` gathers a batch
  with ``gode_qc_collate`` (two CUDA kernels, no host loop, no host->device copy per batch) and returns the same tuple
  with device tensors, index-vector ``E_tgt`` and ``G`` None.  ``DeviceLoader`` iterates (optionally shuffled) batches.
"""
from __future__ import annotations

import numpy as np
import torch

from ... import ops
from ..._lib import check, lib


def _parts(g):
    (M, x, e), o = g
    return np.asarray(M), np.asarray(x, dtype=np.float64), e, np.asarray(o, dtype=np.float64)


def _sorted_edges(e):
    keys = sorted(e.keys())
    src = np.fromiter((k[0] for k in keys), dtype=np.int64, count=len(keys))
    tgt = np.fromiter((k[1] for k in keys), dtype=np.int64, count=len(keys))
    feat = np.asarray([e[k] for k in keys], dtype=np.float64).reshape(len(keys), -1)
    return src, tgt, feat


def collate_g_concat_edge_data(batch, dense=True):
    """QC/datasets/utils.py:153-217.  ``dense=True`` is the reference's return value exactly."""
    parts = [_parts(g) for g in batch]
    n_d, o_d = parts[0][1].shape[1], parts[0][3].shape[0]
    e_d = len(next(iter(parts[0][2].values()))) if len(parts[0][2]) else 0
    ns = np.array([p[0].shape[0] for p in parts], dtype=np.int64)
    ms = np.array([len(p[2]) for p in parts], dtype=np.int64)
    N, M, bs = int(ns.sum()), int(ms.sum()), len(batch)
    n_acc = np.concatenate([[0], np.cumsum(ns)])
    m_acc = np.concatenate([[0], np.cumsum(ms)])
    B = np.repeat(np.arange(bs, dtype=np.int64), ns)
    X = np.zeros([N, n_d])
    E_d = np.zeros([2 * M, e_d])
    E_src = np.zeros([2 * M], dtype=np.int64)
    E_tgt = np.zeros([2 * M], dtype=np.int64)
    Y = np.zeros([bs, o_d])
    G = np.zeros([N, N]) if dense else None
    for b, (Mb, x, e, o) in enumerate(parts):
        n, a, lo, hi = int(ns[b]), int(n_acc[b]), int(m_acc[b]), int(m_acc[b + 1])
        X[a:a + n] = x
        Y[b] = o
        if dense:
            G[a:a + n, a:a + n] = Mb
        if hi > lo:
            src, tgt, feat = _sorted_edges(e)
            # NOTE the reference does not shift the node ids of an edge by n_acc (utils.py:196-203 store `src` / `tgt` as
            # they are in the molecule): its batches are only block-diagonal when every molecule numbers its atoms globally.
            # Kept as is for drop-in parity; MoleculeStore.collate(shift=True) emits the block-diagonal form.
            E_d[lo:hi] = feat
            E_src[lo:hi] = src
            E_tgt[lo:hi] = tgt
            E_d[M + lo:M + hi] = feat
            E_src[M + lo:M + hi] = tgt
            E_tgt[M + lo:M + hi] = src
    Bt, Xt, Edt, Yt = torch.LongTensor(B), torch.FloatTensor(X), torch.FloatTensor(E_d), torch.FloatTensor(Y)
    Es = torch.LongTensor(E_src)
    if not dense:
        return bs, None, Bt, Xt, Edt, Es, torch.LongTensor(E_tgt), Yt
    idx = torch.LongTensor(np.stack([E_tgt, np.arange(2 * M, dtype=np.int64)]))
    Et = torch.sparse_coo_tensor(idx, torch.ones(2 * M), torch.Size([N, 2 * M])).to_dense()
    return bs, torch.FloatTensor(G), Bt, Xt, Edt, Es, Et, Yt


class MoleculeStore:
    """All molecules of a dataset as ragged device arrays (built once; see gode_qc_collate in include/gode.h)."""

    def __init__(self, molecules, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise TypeError("MoleculeStore lives on a CUDA device; graph-odenet_b200 has no CPU path")
        parts = [_parts(g) for g in molecules]
        ns = np.array([p[0].shape[0] for p in parts], dtype=np.int64)
        ms = np.array([len(p[2]) for p in parts], dtype=np.int64)
        self.n_molecules = len(parts)
        self.n_d = parts[0][1].shape[1] if parts else 1
        self.o_d = parts[0][3].shape[0] if parts else 1
        e_d = 0
        srcs, tgts, feats = [], [], []
        for p in parts:
            if len(p[2]):
                s, t, f = _sorted_edges(p[2])
                srcs.append(s), tgts.append(t), feats.append(f)
                e_d = f.shape[1]
        self.e_d = max(e_d, 1)
        cat = lambda xs, w, dt: (np.concatenate(xs) if xs else np.zeros([0, w] if w else [0])).astype(dt)
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.node_ptr = to(np.concatenate([[0], np.cumsum(ns)]).astype(np.int64))
        self.edge_ptr = to(np.concatenate([[0], np.cumsum(ms)]).astype(np.int64))
        self.X_all = to(cat([p[1] for p in parts], self.n_d, np.float32))
        self.Y_all = to(cat([p[3][None] for p in parts], self.o_d, np.float32))
        self.e_src = to(cat(srcs, 0, np.int32))
        self.e_tgt = to(cat(tgts, 0, np.int32))
        self.E_all = to(cat(feats, self.e_d, np.float32))
        self.device = dev

    def collate(self, ids, shift=False):
        """Batch of the molecules ``ids`` (in that order) -> ``(batch_size, None, B, X, E_d, E_src, E_tgt, Y)`` on the device,
        ``E_tgt`` the target-node index of every directed edge.  ``shift=False`` keeps the reference's edge endpoints
        (molecule-local ids, QC/datasets/utils.py:196-203); ``shift=True`` adds the molecule's node offset."""
        dev = self.device
        sel = torch.as_tensor(ids, dtype=torch.int64, device=dev).reshape(-1)
        n_sel = int(sel.numel())
        if n_sel and (int(sel.min()) < 0 or int(sel.max()) >= self.n_molecules):
            raise IndexError("molecule id out of range for a store of %d molecules" % self.n_molecules)
        zero = torch.zeros(1, dtype=torch.int64, device=dev)
        node_off = torch.cat([zero, torch.cumsum(self.node_ptr[sel + 1] - self.node_ptr[sel], 0)])
        edge_off = torch.cat([zero, torch.cumsum(self.edge_ptr[sel + 1] - self.edge_ptr[sel], 0)])
        N, M = (int(v) for v in torch.stack([node_off[-1], edge_off[-1]]).tolist())     # one device -> host read per batch
        B = torch.empty(N, dtype=torch.int64, device=dev)
        X = torch.empty(N, self.n_d, dtype=torch.float32, device=dev)
        E_d = torch.empty(2 * M, self.e_d, dtype=torch.float32, device=dev)
        E_src = torch.empty(2 * M, dtype=torch.int64, device=dev)
        E_tgt = torch.empty(2 * M, dtype=torch.int64, device=dev)
        sel32 = sel.to(torch.int32)
        p = ops._p
        check(lib.gode_qc_collate(n_sel, p(sel32), p(self.node_ptr), p(self.edge_ptr), p(node_off), p(edge_off), int(bool(shift)),
                                  N, M, p(self.X_all), self.n_d, p(self.e_src), p(self.e_tgt), p(self.E_all), self.e_d, p(B), p(X),
                                  p(E_d), p(E_src), p(E_tgt), ops._stream()), "gode_qc_collate")
        return n_sel, None, B, X, E_d, E_src, E_tgt, self.Y_all[sel]


class DeviceLoader:
    """Batches of a MoleculeStore: ``for batch_size, G, B, X, E_d, E_src, E_tgt, Y in loader`` as the reference's
    DataLoader yields them (QC/util.py:127-139), collated on the device."""

    def __init__(self, store, batch_size, shuffle=False, seed=0, shift=True):
        self.store, self.batch_size, self.shuffle, self.shift = store, int(batch_size), shuffle, shift
        self.gen = torch.Generator(device=store.device).manual_seed(seed)

    def __len__(self):
        return (self.store.n_molecules + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = self.store.n_molecules
        order = (torch.randperm(n, device=self.store.device, generator=self.gen) if self.shuffle
                 else torch.arange(n, device=self.store.device))
        for lo in range(0, n, self.batch_size):
            yield self.store.collate(order[lo:lo + self.batch_size], shift=self.shift)
