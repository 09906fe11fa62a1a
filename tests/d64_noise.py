"""Diagnostic (not a test): which tensor of the augmented state drives the dopri5 backward step sequence of the d = 64
(two channels per GroupNorm group) ODE block -- the CPU oracle in float32 and float64, and (with a GPU) the CUDA path.

    python tests/d64_noise.py [d] [gpu]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gcn_ref  # noqa: E402
from tests import _golden as G  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = 512
g = G.load("gcn_golden")
k = "odeblock%d_sub_dopri5/" % d
adj = G.sub_adj()
x0 = G.rnd(40 + d, n, d, scale=0.5)
g1 = G.rnd(50 + d, n, d, scale=1.0 / n)
params = G.params(g, k + "p/")


def show(tag, st):
    tr = st["trace"]
    print("%s: accepted %d rejected %d" % (tag, st.get("accepted", 0), st.get("rejected", 0)))
    worst = np.argmax(np.array([m for _, _, m in tr]), axis=1)
    print("   dominant tensor per step (0 y, 1 a_y, 2 a_t, 3.. a_theta):", np.bincount(worst, minlength=4).tolist())
    for t, dt, m in tr[:12]:
        print("   t %.5f dt %.5f  msr %s" % (t, dt, " ".join("%.2e" % v for v in m)))


for dtype in (torch.float32, torch.float64):
    p = {kk: v.to(dtype).clone().requires_grad_(True) for kk, v in params.items()}
    x = x0.to(dtype).clone().requires_grad_(True)
    st = {}
    y, f = gcn_ref.ode_block(x, adj.to(dtype), p, prefix="odefunc.", method="dopri5", stats=st)
    st["backward"] = {"trace": []}
    y.backward(g1.to(dtype))
    show("oracle %s backward" % str(dtype).split(".")[1], st["backward"])

if len(sys.argv) > 2:
    import graph_odenet_b200  # noqa: F401
    from graph_odenet_b200.GCN import models
    dev = torch.device("cuda:0")
    blk = models.ODEBlock(models.ODEfunc(d), method="dopri5")
    blk.load_state_dict(params)
    blk = blk.to(dev)
    blk.stats = {}
    x = x0.to(dev).requires_grad_(True)
    y = blk(x, adj.to(dev))
    blk.stats["backward"] = {"trace": []}
    y.backward(g1.to(dev))
    show("libgode backward", blk.stats["backward"])

    # One evaluation of the augmented dynamics F(t, (y, a)) = (f, -a^T df/dy, -a^T df/dtheta, -a^T df/dt): the fused kernels
    # and the float32 oracle against the float64 oracle, per tensor, as a fraction of the tensor's max.
    from graph_odenet_b200 import odeint as O, ops

    def oracle_eval(dtype):
        p = {kk: v.to(dtype).clone().requires_grad_(True) for kk, v in params.items()}
        xx = x0.to(dtype).clone().requires_grad_(True)
        tt = torch.tensor(0.37, dtype=dtype, requires_grad=True)
        f = gcn_ref.odefunc(tt, xx, p, adj.to(dtype), "odefunc.")
        keys = ["odefunc.gc1.weight", "odefunc.gc1.bias", "odefunc.norm1.weight", "odefunc.norm1.bias"]
        gs = torch.autograd.grad(f, [xx, tt] + [p[k_] for k_ in keys], -g1.to(dtype))
        return {"f": f.detach(), "a_y": gs[0], "a_t": gs[1].reshape(1), "W": gs[2], "b": gs[3], "gamma": gs[4], "beta": gs[5]}

    r64, r32 = oracle_eval(torch.float64), oracle_eval(torch.float32)
    fm = blk.odefunc
    plan = ops.plan_for(adj.to(dev))
    kern = O._make_kernel(plan, fm.gc1.weight, fm.gc1.bias, fm.norm1.weight, fm.norm1.bias, fm.norm1.num_groups, fm.norm1.eps)
    S, ky, ka, gP = kern.new_S(), kern.new(), kern.new(), kern.new_gP()
    P = kern.n_theta
    gth = torch.empty(P, dtype=torch.float32, device=dev)
    xd, ad = x0.to(dev), g1.to(dev)
    kern.transform(xd, 0.37, S)
    kern.vjp_phase1(S, ad, -1.0, ky, gP)
    kern.vjp_phase2(xd, 0.37, gP, ka, gth)
    nw = (d + 1) * d
    ours = {"f": ky, "a_y": ka, "W": gth[:nw].reshape(d + 1, d), "b": gth[nw:nw + d], "gamma": gth[nw + d:nw + 2 * d],
            "beta": gth[nw + 2 * d:nw + 3 * d], "a_t": gth[P - 1:]}
    print("single evaluation, error vs the float64 oracle as a fraction of the tensor's max:  libgode | float32 oracle")
    for name in ("f", "a_y", "a_t", "W", "b", "gamma", "beta"):
        ref = r64[name]
        sc = float(ref.abs().max())
        e_o = (ours[name].detach().cpu().double().reshape(ref.shape) - ref).abs()
        e_r = (r32[name].double() - ref).abs()
        print("   %-6s max %.2e rms %.2e | max %.2e rms %.2e   (scale %.2e)" % (
            name, float(e_o.max()) / sc, float(e_o.pow(2).mean().sqrt()) / sc, float(e_r.max()) / sc,
            float(e_r.pow(2).mean().sqrt()) / sc, sc))
    if name:
        e = (ours["a_y"].detach().cpu().double() - r64["a_y"]).abs()
        rows = torch.topk(e.max(1).values, 5).indices
        xn = x0.double()
        for r_ in rows.tolist():
            c_ = int(e[r_].argmax())
            pair = xn[r_, (c_ // 2) * 2:(c_ // 2) * 2 + 2] if d == 64 else xn[r_, (c_ // 4) * 4:(c_ // 4) * 4 + 4]
            print("   worst a_y row %d col %d: err %.2e ours %.4e ref %.4e f32-oracle %.4e  group inputs %s" % (
                r_, c_, float(e[r_, c_]), float(ours["a_y"][r_, c_]), float(r64["a_y"][r_, c_]), float(r32["a_y"][r_, c_]),
                pair.tolist()))
