"""Check every gode_gcn_transform call made during an ODEBlock fwd+bwd against a float64 torch evaluation."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graph_odenet_b200  # noqa: E402,F401
from graph_odenet_b200 import odeint, synth  # noqa: E402
from graph_odenet_b200.GCN import models  # noqa: E402

n, d = 4096, 128
dev = torch.device("cuda:0")
torch.manual_seed(0)
row, col, val = synth.powerlaw_graph(n, avg_degree=12, seed=0, device="cpu")
adj = torch.sparse_coo_tensor(torch.stack([row, col]), val, (n, n)).to(dev)
blk = models.ODEBlock(models.ODEfunc(d), method="rk4")
with torch.no_grad():
    blk.odefunc.norm1.weight.uniform_(0.5, 1.5)
    blk.odefunc.norm1.bias.uniform_(-0.5, 0.5)
blk = blk.to(dev)
x = (0.5 * torch.randn(n, d)).to(dev).requires_grad_(True)
g = (torch.randn(n, d) / n).to(dev)

orig = odeint.GcnKernel.transform
calls = []


def checked(self, y, t, out):
    res = orig(self, y, t, out)
    torch.cuda.synchronize()
    z = F.group_norm(y.double(), 32, self.gamma.double(), self.beta.double(), 1e-5)
    ref = z @ self.weight[1:].double() + float(t) * self.weight[0].double()
    err = (res.double() - ref).abs()
    calls.append((float(t), float(err.max() / ref.abs().max()), int((err > 1e-4 * ref.abs().max()).sum()), float(y.abs().max()),
                  y.data_ptr() % 16, out.data_ptr() % 16, bool(torch.isfinite(y).all())))
    return res


odeint.GcnKernel.transform = checked
y = blk(x, adj)
y.backward(g)
for c in calls:
    print("t=%.4f relerr=%.3e bad=%d max|y|=%.3e align=%d/%d finite=%s" % c)
