#!/usr/bin/env python
"""One GAT-ODE rk4 fwd+bwd step of BASELINE config 3 (tools/bench_configs.py config3) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off`; without ncu prints the CUDA-event time of the step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import graph_odenet_b200  # noqa: E402,F401
from graph_odenet_b200 import synth  # noqa: E402
from graph_odenet_b200.GAT import models  # noqa: E402

dev = torch.device("cuda:0")
n = int(os.environ.get("GAT_N", "1000000"))
row, col = synth.powerlaw_graph(n, avg_degree=1 + 2 * 4676 / 3327, seed=0, device=dev, return_raw=True)
keep = row < col
src, tgt = row[keep].contiguous(), col[keep].contiguous()
d, heads = 128, 8
torch.manual_seed(0)
blk = models.ODEBlock(models.ODEfunc(d, heads=heads), method="rk4").to(dev)
x = torch.randn(n, d, device=dev)
g = torch.randn(n, d, device=dev) / n


def step():
    for p in blk.parameters():
        p.grad = None
    xx = x.clone().requires_grad_(True)
    y = blk(xx, src, tgt, None)
    y.backward(g)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
step()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("gat step ms", e0.elapsed_time(e1))
