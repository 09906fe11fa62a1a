"""QC layers with the reference's module surface (QC/layers.py:10-154), computed by libgode kernels.

Same class names, constructor arguments, parameter names / layouts (``MyLinear.weight`` is [in, out]) and
``state_dict`` keys (``mlp.mlp.layers.0.linear.weight`` ...), so checkpoints move between the two.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops


class MyLinear(nn.Module):
    """``mm(input, W) + b`` with W [in, out], U(+-1/sqrt(out)) init (QC/layers.py:10-30) -- gode_linear_f32."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(in_features, out_features))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.weight.size(1))
        with torch.no_grad():
            self.weight.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def forward(self, input, relu=False):
        return ops.LinearFn.apply(input, self.weight, self.bias, relu)


class NonLinear(nn.Module):
    """``f(linear(input))`` (QC/layers.py:33-44); f = relu is fused into the GEMM epilogue."""

    def __init__(self, in_features, out_features, bias=True, f=F.relu):
        super().__init__()
        self.linear = MyLinear(in_features, out_features, bias=bias)
        self.bias = bias
        self.f = f

    def forward(self, input):
        if self.f is F.relu:
            return self.linear(input, relu=True)
        return self.f(self.linear(input))


class MLP(nn.Module):
    """QC/layers.py:46-63 (the reference's ``layer_size`` typo branch -- out_features=None -- raises NameError there;
    here it does what was meant)."""

    def __init__(self, in_features, layer_sizes, out_features=None, bias=True):
        super().__init__()
        if out_features is None:
            out_features = layer_sizes[-1]
            layer_sizes = layer_sizes[:-1]
        layer_inputs = [in_features] + layer_sizes[:-1]
        layers_ = [NonLinear(i, o, bias=bias) for i, o in zip(layer_inputs, layer_sizes)]
        layers_.append(MyLinear(layer_sizes[-1], out_features, bias=bias))
        self.layers = nn.Sequential(*layers_)

    def forward(self, input):
        return self.layers(input)


class TransitionMLP(nn.Module):
    """One hidden ReLU layer of width (in + out) // 2 (QC/layers.py:65-74)."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.mlp = MLP(in_features, [(in_features + out_features) // 2], out_features, bias=bias)

    def forward(self, input):
        return self.mlp(input)


class EdgeEncoderMLP(nn.Module):
    """e [E, fe] -> TransitionMLP(fe -> nf*nf) -> [E, nf, nf] (QC/layers.py:76-86)."""

    def __init__(self, edge_features, node_features, bias=True):
        super().__init__()
        self.mlp = TransitionMLP(edge_features, node_features * node_features, bias=bias)
        self.nf = node_features

    def forward(self, input):
        return self.mlp(input).reshape(input.size(0), self.nf, self.nf)


class EdgeGraphConvolution(nn.Module):
    """``spmm(Etgt, bmm(edge_data, (input W)[Esrc])) + b`` (QC/layers.py:114-154).

    gode_gemm_f32 (input W) -> gode_edge_matvec (per-edge mat-vec, streams edge_data once) ->
    gode_spmm_csr_f32 (segmented sum by target + bias [+ relu]).  ``Etgt``: the reference's dense one-hot [N, E]
    (also accepted: sparse COO, or the target index vector [E])."""

    def __init__(self, in_features, out_features, node_layers=1, edge_layers=1, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(in_features, out_features))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.weight.size(1))
        with torch.no_grad():
            self.weight.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def forward(self, input, Esrc, Etgt, edge_data):
        support = ops.LinearFn.apply(input, self.weight, None, False)
        return ops.edge_message(support, edge_data, Esrc, Etgt, self.bias)

    def extra_repr(self):
        return "%d -> %d" % (self.in_features, self.out_features)
