"""QC models with the reference's names and keyword surface (QC/layer_models.py:27-232; built at
QC/train_egcn.py:118 as ``Model(node_features=, edge_features=, target_features=, hidden_features=, num_layers=,
s2s_processing_steps=, type=, dropout=)`` and called with ``node_features, edge_features, Esrc, Etgt, batch``).

``EdgeODE_K_Sum`` is the builder extension BASELINE config 5 names (the reference maps ``eodesum`` to
``UnimplementedModel``): an ``ODEBlock`` whose function is the reference's ``EdgeGraphConvolution`` wrapped like
``GCN/models.py:161-179`` -- ``f(t, x) = relu(EdgeGC([t || GroupNorm(x)]))``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import odeint as _solver
from .. import ops
from ..GCN.models import GroupNorm
from .layers import EdgeEncoderMLP, EdgeGraphConvolution, MyLinear, TransitionMLP
from .mpnn import MPNN_enn_edge
from .set2set import Set2Set


def get_output_function(type, target_features):
    if type == "regression" or target_features == 1:
        return lambda x: x
    return lambda x: F.log_softmax(x, dim=1)


class UnimplementedModel(nn.Module):
    """QC/layer_models.py:19-24."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("Model not implemented yet")

    def forward(self, *args, **kwargs):
        raise NotImplementedError("Model not implemented yet")


def _n_graphs(batch):
    return int(batch.max().item()) + 1 if batch.numel() else 0      # an empty shard (batch smaller than the world) has no graphs


class EdgeGCN_K_Sum(nn.Module):
    """mlpin -> K x (EdgeGC [+ relu + dropout]) -> mlpout -> scatter_add by molecule (QC/layer_models.py:82-122)."""

    def __init__(self, node_features=None, edge_features=None, target_features=1, hidden_features=73, num_layers=3,
                 s2s_processing_steps=12, type="regression", dropout=0.5, **kwargs):
        super().__init__()
        self.mlpin = TransitionMLP(node_features, hidden_features)
        self.gcmid = nn.ModuleList([EdgeGraphConvolution(hidden_features, hidden_features) for _ in range(num_layers)])
        self.mlpout = TransitionMLP(hidden_features, target_features)
        self.dropout = dropout
        self.ee = EdgeEncoderMLP(edge_features, hidden_features)
        self.type = type
        self.output_function = get_output_function(type, target_features)

    def forward(self, node_features, edge_features, Esrc, Etgt, batch, batch_size=None):
        batch_size = _n_graphs(batch) if batch_size is None else batch_size
        ef = self.ee(edge_features)
        x = self.mlpin(node_features)
        for gc in self.gcmid[:-1]:
            x = F.relu(gc(x, Esrc, Etgt, ef))
            x = F.dropout(x, self.dropout, training=self.training)
        x = self.gcmid[-1](x, Esrc, Etgt, ef)
        x = self.mlpout(x)
        x = ops.scatter_add_rows(x, batch, batch_size)
        return self.output_function(x)


class MPNN_ENN_K_Sum(nn.Module):
    """input Linear -> MPNN_enn_edge(T = num_layers) -> output Linear -> scatter_add (QC/layer_models.py:27-52)."""

    def __init__(self, node_features=None, edge_features=None, target_features=1, hidden_features=73, num_layers=3,
                 s2s_processing_steps=12, type="regression", dropout=0.5, **kwargs):
        super().__init__()
        self.input = nn.Linear(in_features=node_features, out_features=hidden_features)
        self.ee = EdgeEncoderMLP(edge_features, hidden_features)
        self.mpnn = MPNN_enn_edge(edge_features, hidden_features)
        self.mpnn.set_T(num_layers)
        self.output = nn.Linear(in_features=hidden_features, out_features=target_features)
        self.type = type
        self.output_function = get_output_function(type, target_features)

    def forward(self, node_features, edge_features, Esrc, Etgt, batch, batch_size=None):
        batch_size = _n_graphs(batch) if batch_size is None else batch_size
        edge_data = self.ee(edge_features)
        x = ops.LinearFn.apply(node_features, self.input.weight.t(), self.input.bias, False)
        x = self.mpnn(x, Esrc, Etgt, edge_data)
        x = ops.LinearFn.apply(x, self.output.weight.t(), self.output.bias, False)
        x = ops.scatter_add_rows(x, batch, batch_size)
        return self.output_function(x)


class MPNN_ENN_K_Set2Set(nn.Module):
    """QC/layer_models.py:55-82: the message-passing core of ``MPNN_ENN_K_Sum`` (the reference imports ``MPNN_enn_edge as
    MPNN_enn``, :5) with the Set2Set readout, of which the first ``hidden`` columns of q* are kept (:79)."""

    def __init__(self, node_features=None, edge_features=None, target_features=1, hidden_features=73, num_layers=3,
                 s2s_processing_steps=12, type="regression", dropout=0.5, **kwargs):
        super().__init__()
        self.input = nn.Linear(in_features=node_features, out_features=hidden_features)
        self.ee = EdgeEncoderMLP(edge_features, hidden_features)
        self.mpnn = MPNN_enn_edge(edge_features, hidden_features)
        self.mpnn.set_T(num_layers)
        self.s2s = Set2Set(hidden_features, s2s_processing_steps, num_layers=1)
        self.output = nn.Linear(in_features=hidden_features, out_features=target_features)
        self.type = type
        self.output_function = get_output_function(type, target_features)

    def forward(self, node_features, edge_features, Esrc, Etgt, batch, batch_size=None):
        edge_data = self.ee(edge_features)
        x = ops.LinearFn.apply(node_features, self.input.weight.t(), self.input.bias, False)
        x = self.mpnn(x, Esrc, Etgt, edge_data)
        x = self.s2s(x, batch)[:, :x.size(1)]
        x = ops.LinearFn.apply(x, self.output.weight.t(), self.output.bias, False)
        return self.output_function(x)


class EdgeGCN_K_Set2Set(nn.Module):
    """mlpin -> K x (EdgeGC [+ relu + dropout]) -> Set2Set (first ``hidden`` columns of q*) -> mlpout
    (QC/layer_models.py:125-163)."""

    def __init__(self, node_features=None, edge_features=None, target_features=1, hidden_features=73, num_layers=3,
                 s2s_processing_steps=12, type="regression", dropout=0.5, **kwargs):
        super().__init__()
        self.mlpin = TransitionMLP(node_features, hidden_features)
        self.gcmid = nn.ModuleList([EdgeGraphConvolution(hidden_features, hidden_features) for _ in range(num_layers)])
        self.mlpout = TransitionMLP(hidden_features, target_features)
        self.dropout = dropout
        self.ee = EdgeEncoderMLP(edge_features, hidden_features)
        self.s2s = Set2Set(hidden_features, s2s_processing_steps, num_layers=1)
        self.type = type
        self.output_function = get_output_function(type, target_features)

    def forward(self, node_features, edge_features, Esrc, Etgt, batch, batch_size=None):
        ef = self.ee(edge_features)
        x = self.mlpin(node_features)
        for gc in self.gcmid[:-1]:
            x = F.relu(gc(x, Esrc, Etgt, ef))
            x = F.dropout(x, self.dropout, training=self.training)
        x = self.gcmid[-1](x, Esrc, Etgt, ef)
        x = self.s2s(x, batch)[:, :x.size(1)]
        x = self.mlpout(x)
        return self.output_function(x)


class EdgeRES1_K_Set2Set(nn.Module):
    """mlpin -> RESKnorm(hidden, hidden, hidden, nlayers = K, residue_layers = 1) -> Set2Set -> mlpout
    (QC/layer_models.py:166-197).  As in the reference it cannot be built at the default hidden = 73
    (GroupNorm(32, 73) raises ValueError)."""

    def __init__(self, node_features=None, edge_features=None, target_features=1, hidden_features=73, num_layers=3,
                 s2s_processing_steps=12, type="regression", dropout=0.5, **kwargs):
        super().__init__()
        self.mlpin = TransitionMLP(node_features, hidden_features)
        self.gcmid = RESKnorm(hidden_features, hidden_features, hidden_features, nlayers=num_layers, residue_layers=1)
        self.mlpout = TransitionMLP(hidden_features, target_features)
        self.ee = EdgeEncoderMLP(edge_features, hidden_features)
        self.s2s = Set2Set(hidden_features, s2s_processing_steps, num_layers=1)
        self.type = type
        self.output_function = get_output_function(type, target_features)

    def forward(self, node_features, edge_features, Esrc, Etgt, batch, batch_size=None):
        ef = self.ee(edge_features)
        x = self.mlpin(node_features)
        x = self.gcmid(x, Esrc, Etgt, ef)
        x = self.s2s(x, batch)[:, :x.size(1)]
        x = self.mlpout(x)
        return self.output_function(x)


class RESKnorm(nn.Module):
    """QC/layer_models.py:199-232: residual EdgeGC stack with GroupNorm.  (As in the reference, ``zip`` over
    ``gcs[0:-1]`` and ``norms`` stops one layer early, and hidden=73 cannot be built: GroupNorm(32, 73) raises.)"""

    def __init__(self, nfeat, nhid, nclass, nlayers=3, residue_layers=1):
        super().__init__()
        if nlayers < 2 + residue_layers:
            raise ValueError("Can't make a Residual GCN with less than {} layers using {} layers for each residual block"
                             .format(2 + residue_layers, residue_layers))
        self.n_layers = nlayers
        self.gcs = nn.ModuleList([EdgeGraphConvolution(nfeat, nhid)]
                                 + [EdgeGraphConvolution(nhid, nhid) for _ in range(nlayers - 2)]
                                 + [EdgeGraphConvolution(nhid, nclass)])
        self.norms = nn.ModuleList([GroupNorm(min(32, nhid), nhid) for _ in range(nlayers - 2)])
        self.residue_layers = residue_layers

    def forward(self, x, Esrc, Etgt, ef):
        gather_residue, r = 1, None
        for gc, norm in zip(self.gcs[0:-1], self.norms):
            gather_residue -= 1
            if gather_residue == 0:
                r, gather_residue = x, self.residue_layers
            x = norm(F.relu(gc(x, Esrc, Etgt, ef)))
            if gather_residue == 1:
                x = x + r
        if gather_residue > 1:
            x = x + r
        return self.gcs[-1](x, Esrc, Etgt, ef)


# ---------------------------------------------------------------------------------------------
# builder extension: QC edge-conditioned ODE (BASELINE config 5)
# ---------------------------------------------------------------------------------------------


class EdgeODEfunc(nn.Module):
    """f(t, x) = relu(EdgeGC([t || GroupNorm(x)], Esrc, Etgt, ef)) -- GCN/models.py:161-179 with the QC layer.
    The time column enters through the layer's weight [d+1, d] (support = [t || GN(x)] W), so the edge matrices
    stay [E, d, d] as the reference layer requires (square, of the output width)."""

    def __init__(self, dim):
        super().__init__()
        self.norm1 = GroupNorm(_groups(dim), dim)
        self.gc1 = EdgeGraphConvolution(dim + 1, dim)
        self.nfe = 0
        self.graph = None

    def set_adj(self, Esrc, Etgt, ef):
        self.graph = (Esrc, Etgt, ef)

    def forward(self, t, x):
        self.nfe += 1
        xn = self.norm1(x)
        tt = torch.ones_like(xn[:, :1]) * t
        return F.relu(self.gc1(torch.cat([tt, xn], 1), *self.graph))


def _groups(dim):
    """min(32, dim) as GCN/models.py:165 when it divides dim; otherwise the largest divisor <= 32 (hidden = 73 is
    prime: one group)."""
    g = min(32, dim)
    while dim % g:
        g -= 1
    return g


class EdgeODEBlock(nn.Module):
    def __init__(self, odefunc, tol=1e-5, method=None, options=None):
        super().__init__()
        self.odefunc = odefunc
        self.integration_time = torch.tensor([0, 1]).float()
        self.tol, self.method, self.options, self.stats = tol, method, options, None

    def forward(self, x, Esrc, Etgt, ef):
        self.odefunc.set_adj(Esrc, Etgt, ef)
        return _solver.odeint_adjoint_final(self.odefunc, x, self.integration_time, rtol=self.tol, atol=self.tol,
                                            method=self.method, options=self.options, stats=self.stats)

    @property
    def nfe(self):
        return self.odefunc.nfe

    @nfe.setter
    def nfe(self, v):
        self.odefunc.nfe = v


class EdgeODE_K_Sum(nn.Module):
    """mlpin -> ODEBlock(EdgeODEfunc) -> EdgeGC -> mlpout -> scatter_add.  The adjoint treats the edge matrices as
    constants of the ODE function (they are parameters of the dynamics only through ``ee``; their gradient is
    not propagated through the block -- the edge encoder still trains through the trailing EdgeGC)."""

    def __init__(self, node_features=None, edge_features=None, target_features=1, hidden_features=73, num_layers=3,
                 s2s_processing_steps=12, type="regression", dropout=0.5, method=None, options=None, **kwargs):
        super().__init__()
        self.mlpin = TransitionMLP(node_features, hidden_features)
        self.ode = EdgeODEBlock(EdgeODEfunc(hidden_features), method=method, options=options)
        self.gcout = EdgeGraphConvolution(hidden_features, hidden_features)
        self.mlpout = TransitionMLP(hidden_features, target_features)
        self.ee = EdgeEncoderMLP(edge_features, hidden_features)
        self.type = type
        self.output_function = get_output_function(type, target_features)

    @property
    def nfe(self):
        return self.ode.nfe

    @nfe.setter
    def nfe(self, v):
        self.ode.nfe = v

    def forward(self, node_features, edge_features, Esrc, Etgt, batch, batch_size=None):
        batch_size = _n_graphs(batch) if batch_size is None else batch_size
        ef = self.ee(edge_features)
        x = self.mlpin(node_features)
        x = self.ode(x, Esrc, Etgt, ef.detach())
        x = self.gcout(x, Esrc, Etgt, ef)
        x = self.mlpout(x)
        x = ops.scatter_add_rows(x, batch, batch_size)
        return self.output_function(x)
