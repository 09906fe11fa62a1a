"""``Set2Set`` readout with the reference's surface (QC/set2set.py:9-82): ``Set2Set(in_channels, processing_steps,
num_layers=1)``, attribute ``lstm`` (same ``state_dict`` keys), ``forward(x, batch) -> [B, 2 * in_channels]``.

Per processing step the reference runs the LSTM on q*, scores every node against its graph's query, takes a softmax
inside each graph with a Python loop over the batch and boolean masks (:66-70), and scatter-adds a * x (:73).  Here the
scores, the per-graph softmax and the weighted sum are one libgode kernel (one warp per graph, ``gode_segment_attend_*``);
the LSTM recurrence (one time step per processing step) is evaluated on the parameters of ``self.lstm`` with libgode's
fp32 GEMM -- cuDNN's RNN path defaults to TF32 tensor cores (``torch.backends.cudnn.allow_tf32``), which is 1e-4 away from
the reference's fp32 CPU result.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class Set2Set(nn.Module):
    def __init__(self, in_channels, processing_steps, num_layers=1):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = 2 * in_channels
        self.processing_steps = processing_steps
        self.num_layers = num_layers
        self.lstm = nn.LSTM(self.out_channels, self.in_channels, num_layers)
        self.reset_parameters()

    def reset_parameters(self):
        self.lstm.reset_parameters()

    def forward(self, x, batch):
        batch_size = int(batch.max().item()) + 1          # the reference syncs here too (:56)
        batch = batch.to(torch.int64)
        if batch.numel() > 1 and bool((batch[1:] < batch[:-1]).any()):
            order = torch.argsort(batch, stable=True)     # nodes of a graph must be contiguous for the kernel
            x, batch = x[order], batch[order]
        gptr = torch.zeros(batch_size + 1, dtype=torch.int64, device=x.device)
        gptr[1:] = torch.cumsum(torch.bincount(batch, minlength=batch_size), 0)
        gptr = gptr.to(torch.int32)
        h = (x.new_zeros((self.num_layers, batch_size, self.in_channels)),
             x.new_zeros((self.num_layers, batch_size, self.in_channels)))
        q_star = x.new_zeros(batch_size, self.out_channels)
        for _ in range(self.processing_steps):
            if self.num_layers == 1:
                q, h = self._lstm_step(q_star, h)
            else:                                         # stacked LSTM: the library's, in fp32
                with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                    q, h = self.lstm(q_star.unsqueeze(0), h)
                q = q.view(batch_size, self.in_channels)
            r = ops.segment_attend(x, q, gptr)
            q_star = torch.cat([q, r], dim=-1)
        return q_star

    def _lstm_step(self, inp, state):
        """torch.nn.LSTM's recurrence for one time step of a single layer (gate order i, f, g, o)."""
        h, c = state[0][0], state[1][0]
        L = self.lstm
        gates = (ops.LinearFn.apply(inp, L.weight_ih_l0.t(), L.bias_ih_l0, False)
                 + ops.LinearFn.apply(h, L.weight_hh_l0.t(), L.bias_hh_l0, False))
        i, f, g, o = gates.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        return h, (h.unsqueeze(0), c.unsqueeze(0))

    def __repr__(self):
        return "{}({}, {})".format(self.__class__.__name__, self.in_channels, self.out_channels)
