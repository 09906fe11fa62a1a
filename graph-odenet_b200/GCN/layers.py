"""GCN layers with the reference's module surface (GCN/layers.py:9-83), computed by libgode kernels.

Same constructor arguments, attribute names (``weight`` [in, out], ``bias`` [out]), initialisation
(U(-1/sqrt(out), 1/sqrt(out)), GCN/layers.py:25-29) and ``state_dict`` keys, so checkpoints and parameter
tensors move freely between the reference modules and these.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .. import ops


class _GraphConvBase(nn.Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(in_features, out_features))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / math.sqrt(self.weight.size(1))
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def _conv(self, x, adj, relu=False):
        # gode_gemm_f32 (H W) -> gode_spmm_csr_f32 (A_hat . + bias [+ relu]); backward in ops.GraphConvFn
        return ops.graph_conv(x, adj, self.weight, self.bias, relu=relu)

    def extra_repr(self):
        return "%d -> %d" % (self.in_features, self.out_features)


class GraphConvolution(_GraphConvBase):
    """``forward(input, adj) = spmm(adj, mm(input, W)) + b``  (GCN/layers.py:31-37)."""

    def forward(self, input, adj, relu=False):
        return self._conv(input, adj, relu)


class FixedGraphConvolution(_GraphConvBase):
    """Same product with the adjacency held as an attribute so an ODE function has signature f(t, x)
    (GCN/layers.py:45-78)."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__(in_features, out_features, bias)
        self.adj = torch.tensor([[1.0]])

    def set_adj(self, adj):
        self.adj = adj

    def forward(self, input, relu=False):
        return self._conv(input, self.adj, relu)
