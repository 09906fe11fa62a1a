"""Helpers shared by the tests: golden fixture access and reproducible inputs."""
from __future__ import annotations

import os

import numpy as np
import torch

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def load(name):
    if name not in _cache:
        _cache[name] = dict(np.load(os.path.join(HERE, name + ".npz")))
    return _cache[name]


def rnd(seed, *shape, scale=1.0):
    """Same generator as tests/golden/make_golden.py:rnd (numpy legacy RandomState, float32)."""
    return torch.from_numpy((np.random.RandomState(seed).standard_normal(shape) * scale).astype(np.float32))


def fill_params(model, seed=7):
    """Deterministic parameters chosen by NAME (numpy legacy RandomState seeded with crc32(name) ^ seed), so that the fixture
    generator (reference modules) and the tests (B200 modules) hold identical parameters without storing them -- and a
    parameter name that exists on one side only is an immediate KeyError/shape error.  2-D weights are Xavier-normal
    (unit normal for the wide input layers), vectors named ``*norm*.weight`` are 1 + 0.25 N(0,1), every other vector is
    0.1 N(0,1)."""
    import zlib
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            rs = np.random.RandomState((zlib.crc32(name.encode()) ^ seed) & 0x7FFFFFFF)
            v = rs.standard_normal(tuple(p.shape)).astype(np.float32)
            if p.dim() >= 2 and max(p.shape) > 512 and min(p.shape) <= 256:
                pass            # input layers on row-normalised bag-of-words features: unit normal keeps the hidden state O(0.3)
            elif p.dim() >= 2:
                v *= np.float32((2.0 / (p.shape[0] + p.shape[1])) ** 0.5)
                if "odefunc" in name:
                    v *= np.float32(0.5)    # about the reference's own U(+-1/sqrt(out)) scale: keeps the ODE contractive enough to pin
            elif "norm" in name and name.endswith("weight"):
                v = np.float32(1.0) + np.float32(0.25) * v
            else:
                v *= np.float32(0.1)
            p.copy_(torch.from_numpy(v))
    return model


GRAD_SAMPLE = 4096


def grad_sample(t):
    """Large gradient tensors are stored as a fixed strided sample (at most GRAD_SAMPLE elements) plus their norm."""
    flat = torch.as_tensor(t).detach().reshape(-1)
    stride = max(1, (flat.numel() + GRAD_SAMPLE - 1) // GRAD_SAMPLE)
    return flat[::stride]


def pubmed_features(n, nfeat=500, per_row=50, seed=0):
    """``ind.pubmed.allx`` is missing from the reference checkout, so Pubmed runs on the real graph
    with synthetic TF-IDF-like features: ~50 non-zeros per row, row-normalised as GCN/utils.py:185 does."""
    rs = np.random.RandomState(seed)
    x = np.zeros((n, nfeat), np.float32)
    cols = rs.randint(0, nfeat, size=(n, per_row))
    vals = rs.rand(n, per_row).astype(np.float32) + np.float32(0.1)
    np.put_along_axis(x, cols, vals, axis=1)
    x /= x.sum(1, keepdims=True)
    return torch.from_numpy(x)


def params(fix, prefix):
    """state_dict-style mapping of torch tensors stored under ``prefix``."""
    return {k[len(prefix):]: torch.from_numpy(v.copy()) for k, v in fix.items() if k.startswith(prefix)}


def coo_adj(row, col, val, n):
    idx = torch.from_numpy(np.vstack([row, col]).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(val.astype(np.float32)), (n, n))


def cora_adj():
    c = load("planetoid_cora")
    return coo_adj(c["coo_row"], c["coo_col"], c["coo_val"], int(c["n"]))


def sub_adj():
    g = load("gcn_golden")
    return coo_adj(g["sub/row"], g["sub/col"], g["sub/val"], 512)


def dense_features(ds):
    c = load("planetoid_" + ds)
    import scipy.sparse as sp
    m = sp.csr_matrix((c["feat_data"], c["feat_indices"], c["feat_indptr"]), shape=(int(c["n"]), int(c["nfeat"])))
    return torch.from_numpy(np.asarray(m.todense(), dtype=np.float32))


def assert_close(a, b, rtol=1e-5, atol_scale=1e-5, what="", atol_abs=0.0, outliers=None):
    """|a-b| <= rtol*|b| + atol_scale*max|b| + atol_abs elementwise (fp32 parity bar, stated at the call site).

    ``outliers=(frac, tol)``: at most ``frac`` of the elements may miss the bar, and those must still be within
    ``tol * max|b|`` -- for quantities where a few elements are legitimately ill-conditioned (stated at the call site)
    while the bulk must hold the fp32 bar."""
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = float(b.abs().max()) if b.numel() else 0.0
    err = (a - b).abs()
    tol = rtol * b.abs() + atol_scale * scale + atol_abs
    bad = err > tol
    if outliers is not None and bool(bad.any()):
        frac, otol = outliers
        if int(bad.sum()) <= frac * bad.numel() and float(err.max()) <= otol * scale:
            return
    if bool(bad.any()):
        i = int(torch.argmax(err - tol))
        raise AssertionError("%s: max err %.3e (scale %.3e) at flat %d: got %.8e want %.8e; %d/%d bad" % (
            what, float(err.max()), scale, i, float(a.reshape(-1)[i]), float(b.reshape(-1)[i]), int(bad.sum()), bad.numel()))


def assert_close_l2(a, b, tol, what=""):
    """||a-b||_F <= tol * ||b||_F.  Used where single entries are legitimately unstable (gradients of ReLU networks:
    a pre-activation within rounding distance of zero flips its mask; see tools/sensitivity.py)."""
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = float((a - b).norm() / (b.norm() + 1e-300))
    assert err <= tol, "%s: relative L2 error %.3e > %.1e" % (what, err, tol)


def synthetic_molecules(sizes=(6, 1, 9, 3, 2, 7, 4), seed=33, n_d=13, e_d=5, o_d=12, global_ids=False):
    """Dataset items of the reference's QC datasets, ``((M, x, e), o)`` (QC/datasets/qm9.py __getitem__): adjacency matrix,
    node feature rows, ``{(src, tgt): edge feature row}`` with one entry per undirected edge, target row.  A spanning tree
    plus a few extra bonds per molecule; the single-atom molecule has no edge.  Values are float64 with more than 24
    significant bits, so the float32 rounding of the collate is exercised."""
    rng = np.random.RandomState(seed)
    out = []
    for n in sizes:
        pairs = set()
        for i in range(1, n):
            pairs.add((int(rng.randint(0, i)), i))
        for _ in range(n // 3):
            a, b = sorted(int(v) for v in rng.randint(0, n, 2))
            if a != b:
                pairs.add((a, b))
        order = list(pairs)
        rng.shuffle(order)                       # dict insertion order is not the sorted order the collate must produce
        M = np.zeros([n, n])
        e = {}
        for a, b in order:
            if rng.rand() < 0.5:
                a, b = b, a                      # keys are not always (low, high)
            M[a, b] = M[b, a] = 1.0
            e[(a, b)] = list(rng.standard_normal(e_d))
        x = [list(rng.standard_normal(n_d)) for _ in range(n)]
        o = list(rng.standard_normal(o_d))
        out.append(((M, x, e), o))
    return out
