"""Micro-benchmark of the SpMM variants over graph locality (run on the GPU box).

    GODE_SPMM_VARIANT={0,3} GODE_SPMM_BULK={0,1,2} python tools/spmm_sweep.py [N] [d]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graph_odenet_b200  # noqa: E402,F401
from graph_odenet_b200 import ops, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
x = torch.randn(n, d, device=dev)
out = torch.empty_like(x)
for loc, win in ((0.9, 32768), (0.9, 2048), (0.9, 256), (0.9, 64), (0.0, 64), (1.0, 64)):
    row, col, val = synth.powerlaw_graph(n, avg_degree=20, locality=loc, window=win, seed=0, device=dev)
    plan = ops.GraphPlan.from_coo(row, col, val, n, n, build_transpose=False)
    nnz = plan.nnz
    del row, col, val
    for _ in range(3):
        ops.spmm(plan, x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 10
    for _ in range(reps):
        ops.spmm(plan, x, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    comp = nnz * 8 + (n + 1) * 4 + 2 * n * d * 4
    print("variant=%s bulk=%s loc=%.1f win=%6d nnz=%d heavy=%d: %.3f ms  compulsory %.0f GB/s  gather %.0f GB/s" % (
        os.environ.get("GODE_SPMM_VARIANT", "0"), os.environ.get("GODE_SPMM_BULK", "0"), loc, win, nnz, plan.n_heavy, ms,
        comp / ms / 1e6, (nnz * d * 4 + comp) / ms / 1e6), flush=True)
    del plan
