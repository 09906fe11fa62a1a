"""Functional torch-CPU restatement of the reference QC edge-conditioned layers (TEST INFRASTRUCTURE).

Pinned against ``/root/reference/QC/layers.py`` by ``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def edge_graph_convolution(x, esrc, etgt_index, edge_data, weight, bias=None):
    """QC/layers.py:136-149: s = x W; m_e = edge_data[e] @ s[esrc[e]]; out = Etgt m + b.

    ``Etgt`` in the reference is a dense one-hot [N, E] matrix (QC/datasets/utils.py:214); the product
    ``spmm(Etgt, m)`` is restated as ``index_add_`` over the target index vector (identical sums).
    """
    s = torch.mm(x, weight)
    m = torch.bmm(edge_data, s.index_select(0, esrc).unsqueeze(-1)).squeeze(-1)
    out = torch.zeros(x.shape[0], m.shape[1], dtype=x.dtype).index_add_(0, etgt_index, m)
    return out if bias is None else out + bias


def my_linear(x, weight, bias=None):
    """QC/layers.py:26-30: mm(x, W) + b with W stored [in, out]."""
    out = torch.mm(x, weight)
    return out if bias is None else out + bias


def transition_mlp(x, p, prefix):
    """QC/layers.py:65-74: one hidden ReLU layer of width (in+out)//2."""
    h = F.relu(my_linear(x, p[prefix + "mlp.layers.0.linear.weight"], p.get(prefix + "mlp.layers.0.linear.bias")))
    return my_linear(h, p[prefix + "mlp.layers.1.weight"], p.get(prefix + "mlp.layers.1.bias"))


def edge_encoder(e, p, prefix, nf):
    """QC/layers.py:76-86: TransitionMLP(edge_features -> nf*nf) reshaped to [E, nf, nf]."""
    return transition_mlp(e, p, prefix + "mlp.").reshape(e.shape[0], nf, nf)


def edge_gcn_k_sum(node_features, edge_features, esrc, etgt_index, batch, p, num_layers, hidden):
    """QC/layer_models.py:100-122 (EdgeGCN_K_Sum), eval mode, regression output."""
    ef = edge_encoder(edge_features, p, "ee.", hidden)
    x = transition_mlp(node_features, p, "mlpin.")
    for i in range(num_layers - 1):
        x = F.relu(edge_graph_convolution(x, esrc, etgt_index, ef, p["gcmid.%d.weight" % i], p["gcmid.%d.bias" % i]))
    i = num_layers - 1
    x = edge_graph_convolution(x, esrc, etgt_index, ef, p["gcmid.%d.weight" % i], p["gcmid.%d.bias" % i])
    x = transition_mlp(x, p, "mlpout.")
    nb = int(batch.max().item()) + 1
    return torch.zeros(nb, x.shape[1], dtype=x.dtype).index_add_(0, batch, x)


def lstm_step(x, h, c, p, prefix="lstm."):
    """One step of a single-layer torch.nn.LSTM (gate order i, f, g, o; the documented recurrence)."""
    gates = x @ p[prefix + "weight_ih_l0"].t() + p[prefix + "bias_ih_l0"] + h @ p[prefix + "weight_hh_l0"].t() + p[prefix + "bias_hh_l0"]
    i, f, g, o = gates.chunk(4, dim=1)
    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h = torch.sigmoid(o) * torch.tanh(c)
    return h, c


def set2set(x, batch, p, steps, prefix="lstm."):
    """QC/set2set.py:54-77: T x { q = LSTM(q*); e_i = <x_i, q[batch_i]>; a = per-graph softmax(e); r = scatter_add(a x);
    q* = [q || r] }.  Returns q* [B, 2C]."""
    nb = int(batch.max().item()) + 1
    C_ = x.shape[1]
    h = torch.zeros(nb, C_, dtype=x.dtype)
    c = torch.zeros(nb, C_, dtype=x.dtype)
    q_star = torch.zeros(nb, 2 * C_, dtype=x.dtype)
    for _ in range(steps):
        h, c = lstm_step(q_star, h, c, p, prefix)
        q = h
        e = (x * q[batch]).sum(-1)
        a = torch.zeros_like(e)
        for g in range(nb):                                   # the reference's loop over the graphs of the batch (:66-70)
            m = batch == g
            a = a.masked_scatter(m, F.softmax(e[m], dim=0))
        r = torch.zeros(nb, C_, dtype=x.dtype).index_add_(0, batch, a.unsqueeze(1) * x)
        q_star = torch.cat([q, r], dim=-1)
    return q_star


def edge_gcn_k_set2set(node_features, edge_features, esrc, etgt_index, batch, p, num_layers, hidden, steps):
    """QC/layer_models.py:142-163 (EdgeGCN_K_Set2Set), eval mode, regression output."""
    ef = edge_encoder(edge_features, p, "ee.", hidden)
    x = transition_mlp(node_features, p, "mlpin.")
    for i in range(num_layers - 1):
        x = F.relu(edge_graph_convolution(x, esrc, etgt_index, ef, p["gcmid.%d.weight" % i], p["gcmid.%d.bias" % i]))
    i = num_layers - 1
    x = edge_graph_convolution(x, esrc, etgt_index, ef, p["gcmid.%d.weight" % i], p["gcmid.%d.bias" % i])
    x = set2set(x, batch, p, steps, prefix="s2s.lstm.")[:, :hidden]
    return transition_mlp(x, p, "mlpout.")
