"""Kernel-class times of the dense d x d chain at the bench size (transform; A^T gather; VJP dense chain = column sums +
weight gradient + input gradient + GroupNorm backward), from libgode's own CUDA-event profiling scopes.  Environment
switches (GODE_TC_ACC, GODE_WGRAD_RND, ...) are read by the library per launch.  Usage: python tools/dense_chain_time.py [N]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graph_odenet_b200  # noqa: E402,F401
from graph_odenet_b200 import _lib, odeint, ops, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = 128
dev = torch.device("cuda:0")
row, col, val = synth.powerlaw_graph(n, avg_degree=20, seed=0, device=dev)
plan = ops.GraphPlan.from_coo(row, col, val, n, n)
del row, col, val
torch.manual_seed(0)
W = (torch.rand(d + 1, d, device=dev) * 2 - 1) / d ** 0.5
b = (torch.rand(d, device=dev) * 2 - 1) / d ** 0.5
gamma, beta = torch.rand(d, device=dev) + 0.5, torch.rand(d, device=dev) - 0.5
kern = odeint.GcnKernel(plan, W, b, gamma, beta, 32)
y, a = torch.randn(n, d, device=dev), torch.randn(n, d, device=dev)
S, ky, ka, gP = kern.new(), kern.new(), kern.new(), kern.new()
gth = torch.empty(kern.n_theta, device=dev)
lib = _lib.lib


def run(reps=4):
    for _ in range(reps):
        kern.transform(y, 0.3, S)
        kern.vjp_phase1(S, a, 1.0, ky, gP)
        kern.vjp_phase2(y, 0.3, gP, ka, gth)


run(2)
torch.cuda.synchronize()
lib.gode_profile_enable(1)
run(4)
torch.cuda.synchronize()
out = {}
for name, kind in (("agg_fwd", 0), ("agg_t", 1), ("transform", 2), ("vjp_dense", 3)):
    cnt, tot, mx = C.c_int(), C.c_float(), C.c_float()
    lib.gode_profile_read(kind, C.byref(cnt), C.byref(tot), C.byref(mx))
    out[name] = round(tot.value / max(cnt.value, 1), 3)
lib.gode_profile_enable(0)
print({k: os.environ.get(k) for k in ("GODE_TC_ACC", "GODE_WGRAD_RND", "GODE_TC")}, out, flush=True)
