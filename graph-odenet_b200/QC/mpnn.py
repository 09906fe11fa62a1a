"""``MPNN_enn_edge`` (QC/mpnn.py:5-32): T rounds of edge-conditioned messages + a GRU node update."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class MPNN_enn_edge(nn.Module):
    def __init__(self, edge_data_dim, node_data_hidden_dim=200):
        super().__init__()
        self.e_d, self.h_d = edge_data_dim, node_data_hidden_dim
        self.update_net = nn.GRUCell(self.h_d * 2, self.h_d)   # library GEMMs (cuBLAS): [N, 2h] x [2h, 3h], not on the HBM path
        self.T = 8

    def set_T(self, t):
        self.T = t

    def forward(self, x, Esrc, Etgt, edge_data):
        for _ in range(self.T):
            node_msg = ops.edge_message(x, edge_data, Esrc, Etgt)      # QC/mpnn.py:27-29 on libgode
            x = self.update_net(torch.cat([x, node_msg], 1), x)
        return x
