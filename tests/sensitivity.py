"""How sensitive are the adjoint gradients of the smoke problem to a 1e-6 relative perturbation of S?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GODE_TC"] = "0"
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
x0 = (0.5 * torch.randn(n, d)).to(dev)
g = (torch.randn(n, d) / n).to(dev)
orig_t = odeint.GcnKernel.transform
orig_s = odeint.GcnKernel.stage_fwd
NOISE = [0.0]
gen = torch.Generator(device=dev).manual_seed(1)


def noisy(buf):
    if NOISE[0] and buf is not None:
        buf.add_(NOISE[0] * float(buf.abs().max()) * torch.randn(buf.shape, device=dev, generator=gen))


def t_(self, y, t, out):
    r = orig_t(self, y, t, out)
    noisy(out)
    return r


def s_(self, S, k_out, y0=None, kprev=(), coefs=(), coef_self=0.0, y_next=None, t_next=0.0, S_next=None):
    r = orig_s(self, S, k_out, y0, kprev, coefs, coef_self, y_next, t_next, S_next)
    noisy(S_next)
    return r


odeint.GcnKernel.transform = t_
odeint.GcnKernel.stage_fwd = s_


def run(noise):
    NOISE[0] = noise
    for p in blk.parameters():
        p.grad = None
    x = x0.clone().requires_grad_(True)
    y = blk(x, adj)
    y.backward(g)
    f = blk.odefunc
    return {"y": y.detach().clone(), "gx": x.grad.clone(), "gW": f.gc1.weight.grad.clone(), "gb": f.gc1.bias.grad.clone(),
            "gg": f.norm1.weight.grad.clone()}


base = run(0.0)
for nz in (1e-7, 1e-6, 1e-5):
    r = run(nz)
    print("noise %.0e:" % nz, {k: "%.2e" % float((r[k] - base[k]).abs().max() / base[k].abs().max()) for k in base})
# where does grad_x magnitude come from?
gx = base["gx"].abs()
print("grad_x: max %.3e median %.3e, rows with max>100*median: %d" % (gx.max(), gx.median(), int((gx.max(1)[0] > 100 * gx.median()).sum())))
