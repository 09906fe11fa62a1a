"""Kernel timeline of one partitioned fwd+bwd step (torch.profiler / CUPTI; nsys is not in the image).
torchrun --nproc-per-node P tools/trace_step.py [N] -> gpurun_out/trace_rank0.json (chrome trace) and a text summary
of GPU activity per stream: busy time, and how much of the NCCL kernels' time overlaps compute kernels."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graph_odenet_b200  # noqa: E402,F401
from graph_odenet_b200 import parallel, synth  # noqa: E402
from graph_odenet_b200.GCN import models  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = 128
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
parallel.init_process_group(dev)
row, col, val = synth.powerlaw_graph(n, avg_degree=20, locality=0.9, seed=0, device=dev)
plan = parallel.PartitionedPlan.build(row, col, val, n, rank, world)
del row, col, val
torch.cuda.empty_cache()
torch.manual_seed(0)
blk = models.ODEBlock(models.ODEfunc(d), method="rk4").to(dev)
x = torch.randn(plan.n_rows, d, device=dev)


def step():
    for p in blk.parameters():
        p.grad = None
    xx = x.detach().requires_grad_(True)
    y = blk(xx, plan)
    (0.5 * (y * y).sum() / (n * d)).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
dist.barrier()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    out = os.path.join(ROOT, "gpurun_out", "trace_rank0_%dg_%s.json" % (world, plan.mode))
    prof.export_chrome_trace(out)
    ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
    print("mode=%s world=%d span %.2f ms, %d GPU activities" % (plan.mode, world, (t1 - t0) / 1e3, len(ev)))
    by_stream = {}
    for e in ev:
        by_stream.setdefault(e["args"].get("stream"), []).append(e)
    for s, es in by_stream.items():
        busy = sum(e["dur"] for e in es) / 1e3
        names = {}
        for e in es:
            k = e["name"][:50]
            names[k] = names.get(k, 0) + e["dur"] / 1e3
        top = sorted(names.items(), key=lambda kv: -kv[1])[:6]
        print(" stream %s: busy %.2f ms; %s" % (s, busy, "; ".join("%s %.2f" % kv for kv in top)))
    nccl = [e for e in ev if "nccl" in e["name"].lower()]
    comp = [e for e in ev if "nccl" not in e["name"].lower() and e.get("cat") == "kernel" and "gather_rows" not in e["name"]]
    ov = 0.0
    for a in nccl:
        a0, a1 = a["ts"], a["ts"] + a["dur"]
        for c in comp:
            c0, c1 = c["ts"], c["ts"] + c["dur"]
            if c1 > a0 and c0 < a1:
                ov += min(a1, c1) - max(a0, c0)
    print(" nccl kernels: %d, total %.2f ms, overlapped with compute kernels %.2f ms" % (len(nccl), sum(e["dur"] for e in nccl) / 1e3, ov / 1e3))
    # idle gaps on the whole device (no activity on any stream)
    iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev)
    idle, cur = 0.0, iv[0][1]
    for a, b in iv[1:]:
        if a > cur:
            idle += a - cur
        cur = max(cur, b)
    print(" device idle inside the step: %.2f ms" % (idle / 1e3))
    for e in nccl[:14]:
        print("   nccl %-40s start %.2f dur %.3f ms" % (e["name"][:40], (e["ts"] - t0) / 1e3, e["dur"] / 1e3))
dist.destroy_process_group()
