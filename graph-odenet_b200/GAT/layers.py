"""GAT layers with the reference's module surface (GAT/layers.py:11-127), computed by libgode kernels.

Same constructor (``in_features, out_features, bias=True, act=F.relu, eps=1e-6``), the same parameters
(``f = nn.Linear(2*in, out)``, ``w = nn.Linear(2*in, 1)``, Xavier-initialised weights) and ``state_dict`` keys
(``f.weight, f.bias, w.weight, w.bias``), ``forward(x, src, tgt, Mtgt)`` / ``FixedGraphConvolution.set_adj``.
``Mtgt`` (the N x E incidence the reference multiplies by) is accepted and ignored: its CSR -- edges grouped by
target -- is built once from ``tgt`` (ops.GatGraph).

``heads`` is a builder extension (BASELINE config 3): H independent reference heads, outputs
concatenated; ``out_features`` is then the per-head width, ``f`` maps to H*out and ``w`` to H.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops

CHECK_NAN = True   # the reference asserts on NaNs (GAT/layers.py:46-56); costs one device->host read per call


class _GatBase(nn.Module):
    def __init__(self, in_features, out_features, bias=True, act=F.relu, eps=1e-6, heads=1):
        super().__init__()
        if act is not F.relu:
            raise NotImplementedError("the fused GAT kernel applies relu (the reference's default act)")
        self.in_features, self.out_features, self.heads = in_features, out_features, heads
        self.f = nn.Linear(2 * in_features, heads * out_features)
        self.w = nn.Linear(2 * in_features, heads)
        self.eps, self.act = eps, act
        self.reset_parameters()

    def reset_parameters(self):
        if self.heads == 1:
            nn.init.xavier_uniform_(self.f.weight)
            nn.init.xavier_uniform_(self.w.weight)
        else:   # each head initialised as an independent reference layer
            o = self.out_features
            with torch.no_grad():
                for h in range(self.heads):
                    nn.init.xavier_uniform_(self.f.weight[h * o:(h + 1) * o])
                    nn.init.xavier_uniform_(self.w.weight[h:h + 1])

    def _conv(self, x, src, tgt):
        return ops.gat_conv(x, src, tgt, self.f.weight, self.f.bias, self.w.weight, self.w.bias, heads=self.heads,
                            eps=self.eps, check_nan=CHECK_NAN)

    def extra_repr(self):
        return "%d -> %d%s" % (self.in_features, self.out_features, "" if self.heads == 1 else " x %d heads" % self.heads)


class GraphConvolution(_GatBase):
    """GAT/layers.py:31-58.  ``relu`` is accepted for the shared model skeleton: the output is a convex
    combination of relu'd values, so the reference's extra ``F.relu`` around the layer is the identity."""

    def forward(self, x, src, tgt, Mtgt=None, relu=False):
        return self._conv(x, src, tgt)


class FixedGraphConvolution(_GatBase):
    """GAT/layers.py:67-122: the edge list is held as attributes so an ODE function has signature f(t, x)."""

    def __init__(self, in_features, out_features, bias=True, act=F.relu, eps=1e-6, heads=1):
        super().__init__(in_features, out_features, bias, act, eps, heads)
        self.src = self.tgt = self.Mtgt = torch.tensor([[1.0]])

    def set_adj(self, src, tgt, Mtgt=None):
        self.src, self.tgt, self.Mtgt = src, tgt, Mtgt

    def forward(self, x, relu=False):
        return self._conv(x, self.src, self.tgt)
