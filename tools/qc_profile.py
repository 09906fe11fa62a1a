#!/usr/bin/env python
"""One training step of BASELINE config 5 (EdgeGCN_K_Sum, K = 3, hidden 73, 4096 QM9-shaped molecules; the step of
tools/bench_configs.py config5) between cudaProfilerStart/Stop, for `ncu --profile-from-start off`; without ncu prints the
CUDA-event time of the step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import graph_odenet_b200  # noqa: E402,F401
from graph_odenet_b200 import synth  # noqa: E402
from graph_odenet_b200.QC import layer_models  # noqa: E402

dev = torch.device("cuda:0")
n_mol = int(os.environ.get("QC_MOL", "4096"))
b = synth.qm9_like_batch(n_mol, 73, seed=0, device=dev)
torch.manual_seed(0)
model = layer_models.EdgeGCN_K_Sum(node_features=13, edge_features=5, target_features=12, hidden_features=73, num_layers=3).to(dev)
target = torch.randn(n_mol, 12, device=dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)


def step():
    opt.zero_grad(set_to_none=True)
    out = model(b["node_features"], b["edge_features"], b["esrc"], b["etgt"], b["batch"], batch_size=n_mol)
    loss = torch.nn.functional.mse_loss(out, target)
    loss.backward()
    opt.step()


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
print("qc step ms", e0.elapsed_time(e1), "atoms", int(b["node_features"].shape[0]), "edges", int(b["esrc"].numel()))
