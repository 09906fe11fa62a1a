"""Compare the tcgen05 dense kernels with the SIMT path on one ODE-function VJP (run on the GPU box)."""
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.argv = [sys.argv[0], sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]]
    import graph_odenet_b200  # noqa: F401
    from graph_odenet_b200 import odeint, ops, synth
    n, d = int(sys.argv[3]), 128
    scale = float(sys.argv[4])
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    row, col, val = synth.powerlaw_graph(n, avg_degree=12, seed=0, device=dev)
    plan = ops.GraphPlan.from_coo(row, col, val, n, n)
    W = (torch.rand(d + 1, d, device=dev) * 2 - 1) / d ** 0.5
    b = (torch.rand(d, device=dev) * 2 - 1) / d ** 0.5
    gamma = torch.rand(d, device=dev) + 0.5
    beta = torch.rand(d, device=dev) - 0.5
    kern = odeint.GcnKernel(plan, W, b, gamma, beta, 32)
    y = torch.randn(n, d, device=dev)
    a = torch.randn(n, d, device=dev) * scale
    S, ky, ka, gP = kern.new(), kern.new(), kern.new(), kern.new()
    kern.transform(y, 0.3, S)
    gth = torch.empty(kern.n_theta, device=dev)
    kern.vjp_phase1(S, a, 1.0, ky, gP)
    kern.vjp_phase2(y, 0.3, gP, ka, gth)
    torch.cuda.synchronize()
    torch.save({"S": S.cpu(), "ky": ky.cpu(), "ka": ka.cpu(), "gth": gth.cpu()}, sys.argv[2])
    sys.exit(0)

outs = {}
N, SC = sys.argv[1], sys.argv[2]
for tcv in ("0", "1"):
    f = "/tmp/tc_check_%s.pt" % tcv
    env = dict(os.environ, GODE_TC=tcv)
    subprocess.run([sys.executable, __file__, "child", f, N, SC], check=True, env=env)
    outs[tcv] = torch.load(f)
d = 128
print("n", N, "scale", SC)
for k in ("S", "ky", "ka"):
    a, b = outs["0"][k].double(), outs["1"][k].double()
    print("%-4s max|simt|=%.3e  max abs diff=%.3e  rel=%.3e" % (k, a.abs().max(), (a - b).abs().max(), (a - b).abs().max() / a.abs().max()))
g0, g1 = outs["0"]["gth"].double(), outs["1"]["gth"].double()
nw = (d + 1) * d
for name, sl in (("gW0", slice(0, d)), ("gW1", slice(d, nw)), ("gb", slice(nw, nw + d)), ("ggamma", slice(nw + d, nw + 2 * d)),
                 ("gbeta", slice(nw + 2 * d, nw + 3 * d)), ("gt", slice(nw + 3 * d, nw + 3 * d + 1))):
    a, b = g0[sl], g1[sl]
    print("%-6s max|simt|=%.3e  max abs diff=%.3e  rel=%.3e" % (name, a.abs().max(), (a - b).abs().max(), (a - b).abs().max() / a.abs().max()))
