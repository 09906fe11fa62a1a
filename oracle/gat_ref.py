"""Functional torch-CPU restatement of the reference GAT layer (TEST INFRASTRUCTURE, CPU oracle).

Pinned against ``/root/reference/GAT/layers.py`` by ``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def gat_convolution(x, src, tgt, f_w, f_b, w_w, w_b, eps=1e-6, act=F.relu):
    """GAT/layers.py:31-58 (and FixedGraphConvolution :95-122).

    h = [x[src] || x[tgt]];  y = act(f(h));  a = w(h);  a_exp = exp(a - max_all(a));
    out[n] = sum_{e: tgt[e]=n} y_e a_exp_e / (sum_{e: tgt[e]=n} a_exp_e + eps)
    The incidence product ``spmm(Mtgt, .)`` is restated as ``index_add_`` by target (same sums).
    ``f_w`` is [o, 2i] and ``w_w`` [1, 2i] as in ``nn.Linear``.
    """
    n = x.shape[0]
    h = torch.cat([x[src], x[tgt]], dim=1)
    y = act(F.linear(h, f_w, f_b))
    a = F.linear(h, w_w, w_b)
    a_exp = torch.exp(a - torch.max(a, 0, keepdim=True)[0])
    a_sum = torch.zeros(n, 1, dtype=x.dtype).index_add_(0, tgt, a_exp) + eps
    num = torch.zeros(n, y.shape[1], dtype=x.dtype).index_add_(0, tgt, y * a_exp)
    return num / a_sum


def gat_multihead(x, src, tgt, heads, eps=1e-6):
    """Builder extension: H independent reference heads, outputs concatenated."""
    return torch.cat([gat_convolution(x, src, tgt, *h, eps=eps) for h in heads], dim=1)


def gat_odefunc(t, x, p, src, tgt, prefix=""):
    """GAT/models.py:172-179: relu(FixedGC([t || GroupNorm(x)]))."""
    d = x.shape[1]
    xn = F.group_norm(x, min(32, d), p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], 1e-5)
    ttx = torch.cat([torch.ones_like(xn[:, :1]) * t, xn], 1)
    return F.relu(gat_convolution(ttx, src, tgt, p[prefix + "gc1.f.weight"], p[prefix + "gc1.f.bias"],
                                  p[prefix + "gc1.w.weight"], p[prefix + "gc1.w.bias"]))
