"""GAT models (GAT/models.py:8-600): the reference file is GCN/models.py with ``adj`` replaced by
``src, tgt, Mtgt``; the classes here are the GCN skeletons rebound to the GAT layer family."""
from __future__ import annotations

import torch
from torch import nn

from ..GCN import models as _gcn
from ..GCN.models import ODEBlock, GroupNorm, _norm  # noqa: F401  (ODEBlock is family-agnostic)
from .layers import FixedGraphConvolution, GraphConvolution


class ODEfunc(nn.Module):
    """f(t, x) = relu(FixedGC([t*1 || GroupNorm(x)]))  (GAT/models.py:161-179); the trailing relu is the identity."""

    def __init__(self, dim, heads=1):
        super().__init__()
        self.norm1 = _norm(dim)
        if dim % heads:
            raise ValueError("dim must be divisible by heads")
        self.gc1 = FixedGraphConvolution(dim + 1, dim // heads, heads=heads)
        self.nfe = 0

    def set_adj(self, src, tgt, Mtgt=None):
        self.gc1.set_adj(src, tgt, Mtgt)

    def forward(self, t, x):
        self.nfe += 1
        from .. import ops
        from . import layers
        if isinstance(self.norm1, nn.GroupNorm) and ops.gat_ode_fusable(x, self.norm1.num_groups, self.gc1.heads, self.gc1.out_features):
            return ops.gat_ode_func(x, t, self.norm1, self.gc1, check_nan=layers.CHECK_NAN)   # one node, no [t || xn] operand
        xn = self.norm1(x)
        tt = torch.ones_like(xn[:, :1]) * t
        return self.gc1(torch.cat([tt, xn], 1))


class ODEfunc2(nn.Module):
    """GAT/models.py:551-575: two (FixedGC -> relu -> GroupNorm) stages, t prepended to each."""

    def __init__(self, dim, dropout):
        super().__init__()
        self.norm1, self.norm2 = _norm(dim), _norm(dim)
        self.gc1 = FixedGraphConvolution(dim + 1, dim)
        self.gc2 = FixedGraphConvolution(dim + 1, dim)
        self.dropout = dropout
        self.nfe = 0

    def set_adj(self, src, tgt, Mtgt=None):
        self.gc1.set_adj(src, tgt, Mtgt)
        self.gc2.set_adj(src, tgt, Mtgt)

    def forward(self, t, x):
        self.nfe += 1
        tt = torch.ones_like(x[:, :1]) * t
        x = self.norm1(self.gc1(torch.cat([tt, x], 1)))
        return self.norm2(self.gc2(torch.cat([tt, x], 1)))


_gcn.rebind_family(globals(), GraphConvolution, ODEfunc, ODEfunc2)
