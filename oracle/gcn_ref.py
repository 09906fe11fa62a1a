"""Functional torch-CPU restatement of the reference GCN arithmetic (TEST INFRASTRUCTURE, CPU oracle).

Pinned against the reference modules themselves by ``tests/golden/make_golden.py`` (which imports
``/root/reference/GCN/{layers,models}.py`` in the build container and stores their outputs under
``tests/golden/``); ``tests/test_oracle_golden.py`` replays those fixtures through this file.

Everything is a plain function over tensors and a ``state_dict``-keyed parameter mapping, so the same
parameters can be loaded into the reference modules, this oracle and the B200 modules.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import odeint as _ode


def graph_convolution(x, adj, weight, bias=None):
    """GCN/layers.py:31-37 (and FixedGraphConvolution :69-75): ``spmm(adj, mm(x, W)) + b``."""
    out = torch.spmm(adj, torch.mm(x, weight)) if adj.is_sparse else torch.mm(adj, torch.mm(x, weight))
    return out if bias is None else out + bias


def group_norm(x, gamma, beta, eps=1e-5):
    """``nn.GroupNorm(min(32, d), d)`` on an [N, d] input (GCN/models.py:165)."""
    d = x.shape[1]
    return F.group_norm(x, min(32, d), gamma, beta, eps)


def odefunc(t, x, p, adj, prefix=""):
    """GCN/models.py:172-179: ``relu(FixedGC([t*1 || GroupNorm(x)]))``."""
    xn = group_norm(x, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"])
    tt = torch.ones_like(xn[:, :1]) * t
    ttx = torch.cat([tt, xn], 1)
    return F.relu(graph_convolution(ttx, adj, p[prefix + "gc1.weight"], p.get(prefix + "gc1.bias")))


def odefunc2(t, x, p, adj, prefix=""):
    """GCN/models.py:564-575: two (FixedGC -> relu -> GroupNorm) stages with t prepended to each."""
    tt = torch.ones_like(x[:, :1]) * t
    h = F.relu(graph_convolution(torch.cat([tt, x], 1), adj, p[prefix + "gc1.weight"], p.get(prefix + "gc1.bias")))
    h = group_norm(h, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"])
    h = F.relu(graph_convolution(torch.cat([tt, h], 1), adj, p[prefix + "gc2.weight"], p.get(prefix + "gc2.bias")))
    return group_norm(h, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"])


class MaskedOdefunc:
    """``odefunc`` with the ReLU replaced by recorded masks, one per function evaluation in evaluation order:
    ``f = pre * mask`` instead of ``relu(pre) = pre * (pre > 0)``.  Feeding the masks the CUDA path used makes the oracle
    differentiate the SAME piecewise-linear branch, so gradients can be compared at fp32 rounding (1e-5) even when a
    pre-activation within rounding distance of zero would otherwise fall on different sides in the two implementations
    (the parity protocol of SURVEY 8c(5), applied to ReLU masks instead of dropout masks).  ``flips`` counts the elements
    whose recorded mask differs from the oracle's own sign test."""

    def __init__(self, masks):
        self.masks, self.i, self.flips, self.total = list(masks), 0, 0, 0

    def __call__(self, t, x, p, adj, prefix=""):
        xn = group_norm(x, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"])
        tt = torch.ones_like(xn[:, :1]) * t
        pre = graph_convolution(torch.cat([tt, xn], 1), adj, p[prefix + "gc1.weight"], p.get(prefix + "gc1.bias"))
        m = self.masks[self.i]
        self.i += 1
        self.flips += int(((pre.detach() > 0) != m).sum())
        self.total += m.numel()
        return pre * m.to(pre.dtype)


class _Func(torch.nn.Module):
    """Adapter giving a functional ODE function a ``parameters()`` list and an ``nfe`` counter."""

    def __init__(self, fn, p, adj, prefix, keys):
        super().__init__()
        self.fn, self.p, self.adj, self.prefix = fn, p, adj, prefix
        self.keys = keys
        self.nfe = 0

    def parameters(self, recurse=True):  # order = reference ``ODEfunc.parameters()`` order
        return iter([self.p[k] for k in self.keys])

    def forward(self, t, x):
        self.nfe += 1
        return self.fn(t, x, self.p, self.adj, self.prefix)


ODEFUNC_KEYS = ("norm1.weight", "norm1.bias", "gc1.weight", "gc1.bias")
ODEFUNC2_KEYS = ("norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias",
                 "gc1.weight", "gc1.bias", "gc2.weight", "gc2.bias")


def ode_block(x, adj, p, prefix="odefunc.", tol=1e-5, method=None, options=None, stats=None, fn=odefunc,
              keys=ODEFUNC_KEYS, adjoint=True):
    """GCN/models.py:189-193: ``odeint_adjoint(odefunc, x, [0,1], rtol=atol=tol)[1]``.

    Returns ``(y(1), func)``; ``func.nfe`` counts function evaluations as the reference does (:173).
    """
    f = _Func(fn, p, adj, prefix, [prefix + k for k in keys])
    t = torch.tensor([0, 1]).float().type_as(x)
    solver = _ode.odeint_adjoint if adjoint else _ode.odeint
    out = solver(f, x, t, rtol=tol, atol=tol, method=method, options=options, stats=stats)
    return out[1], f


# ------------------------------------------------------------------------------------------
# whole models (GCN/models.py), dropout disabled (eval mode / p=0) -- parity protocol SURVEY 8c(5)
# ------------------------------------------------------------------------------------------


def gcn3(x, adj, p):
    """GCN/models.py:66-81 (GCN3), eval mode."""
    h = F.relu(graph_convolution(x, adj, p["gc1.weight"], p["gc1.bias"]))
    h = F.relu(graph_convolution(h, adj, p["gc2.weight"], p["gc2.bias"]))
    return F.log_softmax(graph_convolution(h, adj, p["gc3.weight"], p["gc3.bias"]), dim=1)


def rgcn3(x, adj, p):
    """GCN/models.py:101-118 (RGCN3), eval mode."""
    h = F.relu(graph_convolution(x, adj, p["gc1.weight"], p["gc1.bias"]))
    h = F.relu(graph_convolution(h, adj, p["gc2.weight"], p["gc2.bias"])) + h
    return F.log_softmax(graph_convolution(h, adj, p["gc3.weight"], p["gc3.bias"]), dim=1)


def odegcn3(x, adj, p, tol=1e-5, method=None, options=None, stats=None):
    """GCN/models.py:204-218 (ODEGCN3), eval mode.  Returns (log-probs, func) for nfe inspection."""
    h = F.relu(graph_convolution(x, adj, p["gc1.weight"], p["gc1.bias"]))
    h, f = ode_block(h, adj, p, prefix="gc2.odefunc.", tol=tol, method=method, options=options, stats=stats)
    return F.log_softmax(graph_convolution(h, adj, p["gc3.weight"], p["gc3.bias"]), dim=1), f
