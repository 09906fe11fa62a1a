"""``MPNN_enn_edge`` (QC/mpnn.py:5-32): T rounds of edge-conditioned messages + a GRU node update."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


def gru_cell(cell, inp, h):
    """``nn.GRUCell`` arithmetic (the reference's ``update_net``, QC/mpnn.py:14,30) with its two products on libgode's
    GEMMs instead of cuBLAS; the gate nonlinearities are elementwise ATen ops (plumbing).  ``cell`` is the ``nn.GRUCell``
    that owns the parameters, so ``state_dict`` keys are the reference's (``update_net.weight_ih`` ...)."""
    gi = ops.LinearFn.apply(inp, cell.weight_ih.t(), cell.bias_ih, False)
    gh = ops.LinearFn.apply(h, cell.weight_hh.t(), cell.bias_hh, False)
    i_r, i_z, i_n = gi.chunk(3, 1)
    h_r, h_z, h_n = gh.chunk(3, 1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1.0 - z) * n + z * h


class MPNN_enn_edge(nn.Module):
    def __init__(self, edge_data_dim, node_data_hidden_dim=200):
        super().__init__()
        self.e_d, self.h_d = edge_data_dim, node_data_hidden_dim
        self.update_net = nn.GRUCell(self.h_d * 2, self.h_d)   # parameter container; the arithmetic is gru_cell() above
        self.T = 8

    def set_T(self, t):
        self.T = t

    def forward(self, x, Esrc, Etgt, edge_data):
        for _ in range(self.T):
            node_msg = ops.edge_message(x, edge_data, Esrc, Etgt)      # QC/mpnn.py:27-29 on libgode
            x = gru_cell(self.update_net, torch.cat([x, node_msg], 1), x)
        return x
