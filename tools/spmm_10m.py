"""SpMM at the bench size (N=10M, ~196M entries, d=128): time of the bare gather and of the gather with the
RK-stage epilogue, per kernel variant.  GODE_SPMM_VARIANT / GODE_SPMM_BULK select the variant (read once per
process).  Usage: python tools/spmm_10m.py [N] [locality] [window] [reps]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graph_odenet_b200  # noqa: E402,F401
from graph_odenet_b200 import _lib, ops, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
loc = float(sys.argv[2]) if len(sys.argv) > 2 else 0.9
win = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
d = 128
dev = torch.device("cuda:0")
row, col, val = synth.powerlaw_graph(n, avg_degree=20, locality=loc, window=win or None, seed=0, device=dev)
plan = ops.GraphPlan.from_coo(row, col, val, n, n, build_transpose=True)
del row, col, val
torch.cuda.empty_cache()
x = torch.randn(n, d, device=dev)
out = torch.empty_like(x)
nnz = plan.nnz
comp = nnz * 8 + (n + 1) * 4 + 2 * n * d * 4


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(lambda: ops.spmm(plan, x, out=out))
print("variant=%s bulk=%s N=%d loc=%.2f win=%d nnz=%d heavy=%d bare: %.3f ms compulsory %.0f GB/s gather-model %.0f GB/s" % (
    os.environ.get("GODE_SPMM_VARIANT", "0"), os.environ.get("GODE_SPMM_BULK", "0"), n, loc, win, nnz, plan.n_heavy, ms,
    comp / ms / 1e6, (nnz * d * 4 + comp) / ms / 1e6), flush=True)

ms_t = timed(lambda: ops.spmm(plan, x, out=out, transpose=True))
print("   A^T gather (values per entry, no row-constant shortcut): %.3f ms" % ms_t, flush=True)
chk = ops.spmm(plan, x)
ref = torch.zeros_like(chk)
# spot check against a segment-sum of 2000 rows (variant-independent reference in fp64)
rows = torch.randint(0, n, (2000,), device=dev)
rp = plan.rowptr.to(torch.int64)
errs = []
for r in rows.tolist()[:200]:
    cols = plan.colidx[rp[r]:rp[r + 1]].to(torch.int64)
    want = (plan.vals[rp[r]:rp[r + 1]].double()[:, None] * x[cols].double()).sum(0)
    errs.append(float((chk[r].double() - want).abs().max() / (want.abs().max() + 1e-30)))
print("   max relative row error vs fp64 on 200 rows: %.2e" % max(errs), flush=True)
del chk, ref

# with the stage-4 epilogue of rk4: bias + relu + y_next = y0 + c1 k1 + c2 k2 + c3 k3 + c_self k
y0 = torch.randn(n, d, device=dev)
ks = [torch.randn(n, d, device=dev) for _ in range(3)]
ynext = torch.empty_like(x)
bias = torch.randn(d, device=dev)
ep = _lib.SpmmEpilogue()
ep.bias = bias.data_ptr()
ep.relu = 1
ep.y0 = y0.data_ptr()
for j, k in enumerate(ks):
    ep.kprev[j] = k.data_ptr()
    ep.coef[j] = 0.125 * (j + 1)
ep.n_prev = 3
ep.coef_self = 0.125
ep.ynext = ynext.data_ptr()
csr = plan.csr(False)
nb = _lib.lib.gode_spmm_workspace_bytes(C.byref(csr), d)
ws = ops.workspace(nb, dev, "spmm") if nb else None


def full():
    _lib.check(_lib.lib.gode_spmm_csr_f32(C.byref(csr), ops._p(x), d, d, ops._p(out), d, C.byref(ep), ops._p(ws), nb,
                                          ops._stream()), "spmm")


ms2 = timed(full)
print("   with stage-4 epilogue (6 extra [N,d] streams = %.1f GB): %.3f ms" % (6 * n * d * 4 / 1e9, ms2), flush=True)
# pure streaming reference point: out = x (copy) on the same tensors
ms3 = timed(lambda: out.copy_(x))
print("   copy [N,d]: %.3f ms = %.0f GB/s" % (ms3, 2 * n * d * 4 / ms3 / 1e6), flush=True)
