"""Isolate regressions: ODEBlock rk4 fwd+bwd under kernel-selection env vars, against the CPU oracle."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import graph_odenet_b200  # noqa: F401
    from graph_odenet_b200 import synth
    from graph_odenet_b200.GCN import models
    n, d = 4096, 128
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    row, col, val = synth.powerlaw_graph(n, avg_degree=12, seed=0, device="cpu")
    adj = torch.sparse_coo_tensor(torch.stack([row, col]), val, (n, n)).to(dev)
    blk = models.ODEBlock(models.ODEfunc(d), method="rk4")
    with torch.no_grad():
        blk.odefunc.norm1.weight.uniform_(0.5, 1.5)
        blk.odefunc.norm1.bias.uniform_(-0.5, 0.5)
    x_cpu = 0.5 * torch.randn(n, d)
    g_cpu = torch.randn(n, d) / n
    out = {}
    if sys.argv[3] == "oracle":
        from oracle import gcn_ref
        p = {k: v.detach().clone().requires_grad_(True) for k, v in blk.state_dict().items()}
        xo = x_cpu.clone().requires_grad_(True)
        yo, _ = gcn_ref.ode_block(xo, adj.cpu(), p, prefix="odefunc.", method="rk4")
        yo.backward(g_cpu)
        out = {"y": yo.detach(), "gx": xo.grad, "gW": p["odefunc.gc1.weight"].grad, "gb": p["odefunc.gc1.bias"].grad,
               "gg": p["odefunc.norm1.weight"].grad, "gbeta": p["odefunc.norm1.bias"].grad}
    else:
        blk = blk.to(dev)
        x = x_cpu.to(dev).requires_grad_(True)
        y = blk(x, adj)
        y.backward(g_cpu.to(dev))
        f = blk.odefunc
        out = {"y": y.detach().cpu(), "gx": x.grad.cpu(), "gW": f.gc1.weight.grad.cpu(), "gb": f.gc1.bias.grad.cpu(),
               "gg": f.norm1.weight.grad.cpu(), "gbeta": f.norm1.bias.grad.cpu()}
    torch.save(out, sys.argv[2])
    sys.exit(0)

runs = {}
for name, env in (("oracle", {}), ("tc0", {"GODE_TC": "0"}), ("tc1_transform", {"GODE_TC": "1"}),
                  ("tc2_igrad", {"GODE_TC": "2"}), ("tc4_wgrad", {"GODE_TC": "4"}), ("tc7", {"GODE_TC": "7"})):
    f = "/tmp/adj_%s.pt" % name
    subprocess.run([sys.executable, __file__, "child", f, "oracle" if name == "oracle" else "gpu"], check=True,
                   env=dict(os.environ, **env))
    runs[name] = torch.load(f)
ref = runs["oracle"]
for name, r in runs.items():
    if name == "oracle":
        continue
    print(name, {k: "%.2e" % float((r[k].double() - ref[k].double()).abs().max() / ref[k].double().abs().max()) for k in ref})
