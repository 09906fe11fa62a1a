"""Seeded synthetic workloads of the shapes BASELINE.json names (no datasets can be downloaded here).

``powerlaw_graph`` -- config 4: an undirected power-law graph returned as the reference loader would
return it: ``normalize(A + I)`` (GCN/utils.py:186,205-212) as an int64 COO with fp32 values 1/(deg+1).

Degrees follow a Chung-Lu model with weights w_i ~ rank^(-1/(exponent-1)) (hubs are scattered over the id
range by a seeded permutation).  ``locality`` is the fraction of edges whose second endpoint is drawn near
the first in id space (two-sided geometric offset of scale ``window``): it models a graph whose ids have
been through a locality-preserving reordering (RCM / Rabbit order -- a pure relabelling, parity-neutral),
which is what makes both the L2-resident gather and a small multi-GPU halo possible.  ``locality=0`` is the
adversarial fully random Chung-Lu graph.  Everything is torch ops, so it runs on the GPU at 2e8 edges in
well under a second and on the CPU for tests.
"""
from __future__ import annotations

import torch


def powerlaw_graph(n, avg_degree=20, exponent=2.2, max_degree=100_000, locality=0.9, window=None, seed=0,
                   device="cuda", return_raw=False):
    """Returns (row, col, val) of D^-1 (A + I): int64, int64, float32; sorted by (row, col)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    m = max(int(n * (avg_degree - 1) / 2), 1)          # undirected edges to draw (self-loops add n entries)
    window = window or max(64, min(n // 8, 32768))
    rank = torch.arange(1, n + 1, device=dev, dtype=torch.float64)
    w = rank.pow(-1.0 / (exponent - 1.0))
    w = torch.clamp(w * (2.0 * m / w.sum()), max=float(max_degree))   # expected degree, clipped
    perm = torch.randperm(n, generator=gen, device=dev)
    w = w[perm]
    cdf = torch.cumsum(w, 0)
    total = cdf[-1]

    def draw(k):
        u = torch.rand(k, generator=gen, device=dev, dtype=torch.float64) * total
        return torch.searchsorted(cdf, u).clamp_(max=n - 1)

    src = draw(m)
    dst = draw(m)
    if locality > 0:
        local = torch.rand(m, generator=gen, device=dev) < locality
        # two-sided geometric offset, |offset| >= 1
        u = torch.rand(m, generator=gen, device=dev, dtype=torch.float64).clamp_(min=1e-12)
        mag = (-(u.log()) * (window / 3.0)).floor().to(torch.int64) + 1
        sign = torch.where(torch.rand(m, generator=gen, device=dev) < 0.5, -1, 1)
        near = src + sign * mag
        near = torch.where(near < 0, -near, near)
        near = torch.where(near >= n, 2 * (n - 1) - near, near).clamp_(0, n - 1)
        dst = torch.where(local, near, dst)
        del local, u, mag, sign, near
    keep = src != dst
    src, dst = src[keep], dst[keep]
    lo, hi = torch.minimum(src, dst), torch.maximum(src, dst)
    key = torch.unique(lo * n + hi)                    # undirected simple graph (nx.Graph collapses duplicates)
    lo, hi = key // n, key % n
    del key, src, dst, keep
    diag = torch.arange(n, device=dev, dtype=torch.int64)
    row = torch.cat([lo, hi, diag])
    col = torch.cat([hi, lo, diag])
    del lo, hi
    order = torch.argsort(row * n + col)
    row, col = row[order], col[order]
    del order
    if return_raw:
        return row, col
    deg = torch.bincount(row, minlength=n).to(torch.float64)      # row sums of A + I
    r_inv = 1.0 / deg
    r_inv[torch.isinf(r_inv)] = 0.0
    val = r_inv[row].to(torch.float32)
    return row, col, val


def qm9_like_batch(n_mol, hidden, seed=0, device="cuda", edge_feat=5, node_feat=13):
    """Config 5: block-diagonal batch of molecule-shaped graphs (about 18 atoms, tree + ring closures).

    Returns dict(node_features [N,13], edge_features [E,5], esrc [E], etgt [E], batch [N]) with node ids
    offset per molecule (the correct block-diagonal form).
    """
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    atoms = (torch.randn(n_mol, generator=gen, device=dev) * 3 + 18).round().clamp_(3, 29).to(torch.int64)
    offs = torch.cumsum(atoms, 0) - atoms
    N = int(atoms.sum().item())
    batch = torch.repeat_interleave(torch.arange(n_mol, device=dev), atoms)
    local = torch.arange(N, device=dev) - offs[batch]
    # tree: atom i>0 bonds to a random earlier atom of its molecule
    child = torch.nonzero(local > 0).squeeze(1)
    parent = offs[batch[child]] + (torch.rand(child.numel(), generator=gen, device=dev) * local[child]).floor().to(torch.int64)
    # one ring closure per molecule between two random distinct atoms
    a = offs + (torch.rand(n_mol, generator=gen, device=dev) * atoms).floor().to(torch.int64)
    b = offs + (torch.rand(n_mol, generator=gen, device=dev) * atoms).floor().to(torch.int64)
    ok = a != b
    u = torch.cat([child, a[ok]])
    v = torch.cat([parent, b[ok]])
    esrc = torch.cat([u, v])
    etgt = torch.cat([v, u])
    E = esrc.numel()
    nf = torch.zeros(N, node_feat, device=dev)
    kind = torch.randint(0, 5, (N,), generator=gen, device=dev)
    nf[torch.arange(N, device=dev), kind] = 1.0
    nf[:, 5:] = torch.rand(N, node_feat - 5, generator=gen, device=dev)
    ef = torch.zeros(E, edge_feat, device=dev)
    half = E // 2
    dist = 1.0 + 0.6 * torch.rand(half, generator=gen, device=dev)
    bond = torch.randint(0, 4, (half,), generator=gen, device=dev)
    ef[:half, 0] = dist
    ef[half:, 0] = dist
    ef[torch.arange(half, device=dev), 1 + bond] = 1.0
    ef[half + torch.arange(half, device=dev), 1 + bond] = 1.0
    return {"node_features": nf, "edge_features": ef, "esrc": esrc, "etgt": etgt, "batch": batch, "n_mol": n_mol}
