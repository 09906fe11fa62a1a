"""numpy restatement of the reference's graph preprocessing (TEST INFRASTRUCTURE, CPU oracle).

Integer / index work: every function here is the *bit-exact* target for the device-side plan builder.

reference                                   here
-----------------------------------------   -------------------------------------------
GCN/utils.py:180  nx.adjacency_matrix(...)  adjacency_from_dict_of_lists
GCN/utils.py:186  normalize(adj + I)        add_self_loops + normalize_rows
GCN/utils.py:205-212 normalize              normalize_rows  (float64 row sums, inf -> 0)
GCN/utils.py:222-229 scipy COO -> torch COO to_coo_f32      (float32 cast, int64 indices)
(torch.spmm consumes the COO as is)         coo_to_csr / csr_transpose  (canonical form used by the kernels)
-- builder extension --         partition_rows / local_csr  (1-D row partition + halo renumbering)
"""
from __future__ import annotations

import numpy as np


def adjacency_from_dict_of_lists(graph):
    """Undirected simple adjacency (COO, both directions, self-loops once) of ``nx.from_dict_of_lists``.

    Node numbering follows networkx: dictionary keys in insertion order first, then unseen neighbours in
    encounter order.  Duplicate neighbour entries collapse (nx.Graph).  Returns (n, row, col) with entries
    sorted by (row, col) -- the order ``nx.adjacency_matrix`` (CSR) yields.
    """
    index = {}
    for u in graph:
        index.setdefault(u, len(index))
    for u, nbrs in graph.items():
        for v in nbrs:
            index.setdefault(v, len(index))
    n = len(index)
    pairs = set()
    for u, nbrs in graph.items():
        iu = index[u]
        for v in nbrs:
            iv = index[v]
            pairs.add((iu, iv))
            pairs.add((iv, iu))
    arr = np.array(sorted(pairs), dtype=np.int64).reshape(-1, 2)
    return n, arr[:, 0].copy(), arr[:, 1].copy()


def add_self_loops(n, row, col, val):
    """``adj + sp.eye(n)``: existing diagonal entries are incremented, missing ones created.  float64."""
    key = row * n + col
    dkey = np.arange(n, dtype=np.int64) * (n + 1)
    allk = np.concatenate([key, dkey])
    allv = np.concatenate([val.astype(np.float64), np.ones(n)])
    uk, inv = np.unique(allk, return_inverse=True)
    uv = np.zeros(len(uk))
    np.add.at(uv, inv, allv)
    return uk // n, uk % n, uv


def normalize_rows(n, row, col, val):
    """GCN/utils.py:205-212: D^-1 * M with float64 row sums, 1/0 -> 0."""
    rowsum = np.zeros(n)
    np.add.at(rowsum, row, val.astype(np.float64))
    with np.errstate(divide="ignore"):
        r_inv = np.power(rowsum, -1.0)
    r_inv[np.isinf(r_inv)] = 0.0
    return r_inv[row] * val.astype(np.float64)


def to_coo_f32(row, col, val):
    """GCN/utils.py:222-229: int64 indices, float32 values."""
    return row.astype(np.int64), col.astype(np.int64), val.astype(np.float32)


def coo_to_csr(n_rows, row, col, val):
    """Canonical CSR of a (possibly uncoalesced, unsorted) COO.

    Entries are ordered by (row, col) with a stable sort; entries with equal (row, col) are summed in
    float32 in their original order (what a sequential COO spmm does implicitly).
    Returns int32 rowptr [n_rows+1], int32 colidx [nnz'], float32 vals [nnz'].
    """
    row = np.asarray(row, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    val = np.asarray(val, dtype=np.float32)
    order = np.lexsort((col, row))  # stable: primary row, secondary col, ties keep input order
    r, c, v = row[order], col[order], val[order]
    if len(r):
        first = np.ones(len(r), dtype=bool)
        first[1:] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
    else:
        first = np.zeros(0, dtype=bool)
    starts = np.flatnonzero(first)
    out_v = v[starts].copy()
    if len(starts) != len(r):  # sequential float32 accumulation of duplicates
        seg = np.cumsum(first) - 1
        dup = np.flatnonzero(~first)
        for i in dup:  # rare path; order preserved
            out_v[seg[i]] = np.float32(out_v[seg[i]] + v[i])
    out_r, out_c = r[starts], c[starts]
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(rowptr, out_r + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr.astype(np.int32), out_c.astype(np.int32), out_v.astype(np.float32)


def csr_transpose(n_rows, n_cols, rowptr, colidx, vals):
    """CSR of the transpose; entries ordered by (col, row).  Also returns the permutation (T entry -> entry)."""
    rows = np.repeat(np.arange(n_rows, dtype=np.int64), np.diff(rowptr.astype(np.int64)))
    order = np.argsort(colidx.astype(np.int64), kind="stable")
    rowptr_t = np.zeros(n_cols + 1, dtype=np.int64)
    np.add.at(rowptr_t, colidx.astype(np.int64) + 1, 1)
    rowptr_t = np.cumsum(rowptr_t)
    return rowptr_t.astype(np.int32), rows[order].astype(np.int32), vals[order].astype(np.float32), order.astype(np.int32)


def csr_spmm(rowptr, colidx, vals, x):
    """Plain sequential float32 CSR SpMM (small cases only) -- order of accumulation = entry order."""
    n = len(rowptr) - 1
    y = np.zeros((n, x.shape[1]), dtype=np.float32)
    for r in range(n):
        acc = np.zeros(x.shape[1], dtype=np.float32)
        for e in range(rowptr[r], rowptr[r + 1]):
            acc = acc + vals[e] * x[colidx[e]]
        y[r] = acc
    return y


def partition_rows(n, world):
    """Contiguous 1-D row partition: rank p owns [bounds[p], bounds[p+1])."""
    return np.array([(p * n) // world for p in range(world + 1)], dtype=np.int64)


def local_csr(rowptr, colidx, vals, bounds, rank):
    """Row block of ``rank`` with columns renumbered to [owned | halo].

    halo = sorted unique global column ids outside the owned range.  Returns a dict with
    ``rowptr, colidx, vals`` (local), ``halo`` (global ids, int64, sorted), ``halo_owner_ptr``
    (halo[halo_owner_ptr[q]:halo_owner_ptr[q+1]] is owned by rank q).
    """
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    e0, e1 = int(rowptr[lo]), int(rowptr[hi])
    cols = colidx[e0:e1].astype(np.int64)
    owned = (cols >= lo) & (cols < hi)
    halo = np.unique(cols[~owned])
    local = np.where(owned, cols - lo, 0)
    if len(halo):
        local[~owned] = (hi - lo) + np.searchsorted(halo, cols[~owned])
    owner_ptr = np.searchsorted(halo, bounds)
    return {
        "rowptr": (rowptr[lo:hi + 1].astype(np.int64) - e0).astype(np.int32),
        "colidx": local.astype(np.int32),
        "vals": vals[e0:e1].copy(),
        "halo": halo,
        "halo_owner_ptr": owner_ptr.astype(np.int64),
        "lo": lo,
        "hi": hi,
    }
