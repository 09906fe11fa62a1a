"""One SpMM configuration, a few launches (target for `ncu --set full`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graph_odenet_b200  # noqa: E402,F401
from graph_odenet_b200 import ops, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
win = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
loc = float(sys.argv[3]) if len(sys.argv) > 3 else 0.9
d = 128
dev = torch.device("cuda:0")
x = torch.randn(n, d, device=dev)
out = torch.empty_like(x)
row, col, val = synth.powerlaw_graph(n, avg_degree=20, locality=loc, window=win, seed=0, device=dev)
plan = ops.GraphPlan.from_coo(row, col, val, n, n, build_transpose=False)
for _ in range(5):
    ops.spmm(plan, x, out=out)
torch.cuda.synchronize()
print("ok", plan.nnz)
