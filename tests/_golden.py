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


def assert_close(a, b, rtol=1e-5, atol_scale=1e-5, what="", atol_abs=0.0):
    """|a-b| <= rtol*|b| + atol_scale*max|b| + atol_abs elementwise (fp32 parity bar, stated at the call site)."""
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = float(b.abs().max()) if b.numel() else 0.0
    err = (a - b).abs()
    tol = rtol * b.abs() + atol_scale * scale + atol_abs
    bad = err > tol
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
