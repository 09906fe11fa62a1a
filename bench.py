#!/usr/bin/env python
"""Benchmark of the hot path: GCN-ODE (one ODE block, RK4 3/8-rule, t in [0,1]) forward + adjoint backward
on a synthetic power-law graph -- BASELINE.json's metric ("ODE func-evals/sec and edges/sec ... HBM GB/s").

    python bench.py --gpus N --steps K --warmup W            # ours (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference arithmetic on the host CPU (oracle port)

One JSON line on stdout (rank 0).  A *step* = ODEBlock forward + loss + adjoint backward + optimiser step on
the ODE function's parameters = 9 function evaluations (4 forward, 1 + 4 in the adjoint).
``value`` = nnz(A_hat) * func-evals / time  [edges/s], inputs resident in HBM; ``e2e`` = the same metric
through the public module API with the step's input copied from pinned host memory and the loss read back.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "gcn_ode_rk4_fwd_bwd_edges_per_sec"
UNIT = "edges/s"
CPU_SAMPLE_NODES = 1_000_000   # the CPU legs run the workload's generator at N = 1 M / nnz = 20 M in full
NFE_PER_STEP = 9  # rk4 on t=[0,1]: 4 forward + (1 + 4) adjoint evaluations of ODEfunc (GCN/models.py:173 counter)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--nodes", type=int, default=10_000_000)
    ap.add_argument("--avg-degree", type=float, default=20.0)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--locality", type=float, default=0.9)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--method", default="rk4")
    ap.add_argument("--cpu-sample-nodes", type=int, default=CPU_SAMPLE_NODES,
                    help="nodes of the CPU sample of the workload (same generator); both CPU legs use the same N")
    ap.add_argument("--cpu-budget-s", type=float, default=200.0, help="--impl reference: wall-clock budget of the timed steps")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--halo-mode", default=None, help="N>1: sync | async | split (NCCL) | p2p | p2p-async (peer memory); "
                                                      "default: GODE_HALO_MODE or the library default")
    ap.add_argument("--shuffle-ids", action="store_true", help="relabel the generated graph with a random permutation first "
                                                               "(a graph whose locality is hidden from the id order)")
    ap.add_argument("--reorder", choices=["none", "degree", "rcm"], default="none",
                    help="locality-aware relabelling pass (parallel.locality_order) applied before the plan is built")
    ap.add_argument("--reorder-sweeps", type=int, default=0, help="barycentre sweeps after --reorder")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(a):
    extra = (", ids shuffled" if getattr(a, "shuffle_ids", False) else "") + (
        ", relabelled by %s%s" % (a.reorder, " + %d barycentre sweeps" % a.reorder_sweeps if a.reorder_sweeps else "")
        if getattr(a, "reorder", "none") != "none" else "")
    return "gcn-ode rk4 fwd+bwd, synthetic power-law graph N=%d avg_deg=%g d=%d locality=%g seed=%d%s" % (
        a.nodes, a.avg_degree, a.dim, a.locality, a.seed, extra)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons.  Started before the warm-up (nvidia-smi takes a few hundred ms to
    produce its first row); ``mark_begin`` / ``mark_end`` bracket the timed region and the summary uses the rows that
    fall inside it -- or, when the region is shorter than the sampling period, the rows of the warm-up that ran the
    same workload right before it (said so in ``window``)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, period_ms=100):
        self.rows, self.proc, self.index, self.period_ms = [], None, index, period_ms
        self.t_load = self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_load(self):
        self.t_load = time.perf_counter()

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc:
            time.sleep(self.period_ms / 1e3)      # let the row of the last interval arrive
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass

    def summary(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def collect(lo, hi):
            sm, mx, reasons = [], [], set()
            for ts, r in self.rows:
                if lo is not None and not (lo <= ts <= hi):
                    continue
                try:
                    sm.append(float(r[0]))
                    mx.append(float(r[1]))
                    for nm, v in zip(names, r[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
                except Exception:
                    continue
            return sm, mx, reasons

        window = "timed region"
        sm, mx, reasons = collect(self.t0, (self.t1 or 0) + 2 * self.period_ms / 1e3) if self.t0 is not None else ([], [], set())
        if len(sm) < 2 and self.t_load is not None:
            window = "warm-up + timed region (same workload; the timed region is shorter than two sampling periods)"
            sm, mx, reasons = collect(self.t_load, (self.t1 or time.perf_counter()) + 2 * self.period_ms / 1e3)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": "no nvidia-smi rows"}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference arithmetic on the host cores
# --------------------------------------------------------------------------------------------------


def _load_synth():
    """The workload generator, loaded by path: importing the package would dlopen libgode.so, and the CPU legs must not
    map the product's native code."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_gode_synth", os.path.join(ROOT, "graph-odenet_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_reference_step(n, a, threads):
    """Builds the sample problem and returns a closure running one fwd+bwd step with the reference arithmetic
    (torch.spmm on the COO tensor + torch.mm + GroupNorm, restated solver) -- oracle/gcn_ref.py."""
    import torch
    synth = _load_synth()
    from oracle import gcn_ref
    torch.set_num_threads(threads)
    row, col, val = synth.powerlaw_graph(n, avg_degree=a.avg_degree, locality=a.locality, seed=a.seed, device="cpu")
    nnz = int(val.numel())
    adj = torch.sparse_coo_tensor(torch.stack([row, col]), val, (n, n))
    d = a.dim
    g = torch.Generator().manual_seed(a.seed)
    bound = 1.0 / d ** 0.5
    p = {"odefunc.norm1.weight": torch.ones(d), "odefunc.norm1.bias": torch.zeros(d),
         "odefunc.gc1.weight": (torch.rand(d + 1, d, generator=g) * 2 - 1) * bound,
         "odefunc.gc1.bias": (torch.rand(d, generator=g) * 2 - 1) * bound}
    p = {k: v.requires_grad_(True) for k, v in p.items()}
    x = torch.randn(n, d, generator=g)

    def step():
        for v in p.values():
            v.grad = None
        xx = x.clone().requires_grad_(True)
        y, f = gcn_ref.ode_block(xx, adj, p, prefix="odefunc.", method=a.method)
        loss = 0.5 * (y * y).mean()
        loss.backward()
        return f.nfe, float(loss)

    return step, nnz


def run_cpu_sample(a, steps, warmup, budget_s=None):
    """The reference arithmetic (oracle port) on the host cores, on the workload's generator at N = --cpu-sample-nodes.
    Both CPU legs (``cpu_baseline`` of the default run: 1 timed step; ``--impl reference``: up to K timed steps inside
    ``budget_s``) use the SAME sample, so their edges/s are directly comparable."""
    import torch
    threads = os.cpu_count() or 1
    n = min(a.cpu_sample_nodes, a.nodes)
    step, nnz = cpu_reference_step(n, a, threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    nfe = done = 0
    for _ in range(steps):
        k, _ = step()
        nfe += k
        done += 1
        el = time.perf_counter() - t0
        if budget_s is not None and el + el / done > budget_s:     # the next step would not fit the budget
            break
    dt = time.perf_counter() - t0
    value = nnz * nfe / dt
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "steps_timed": done,
            "sample": "same generator at N=%d nnz=%d d=%d (SURVEY 8d: the 1 M-node instance of the workload in full), %d warm-up + %d "
                      "timed rk4 fwd+bwd steps (%.1f s), torch %s CPU with %d threads, torch.spmm on the COO tensor as the reference builds it"
                      % (n, nnz, a.dim, warmup, done, dt, torch.__version__, threads),
            "ms_per_step": dt / max(done, 1) * 1e3, "func_evals_per_sec": nfe / dt}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # one warm-up step (a step is ~30 s of CPU work at N = 1 M) and as many of the K requested steps as fit the budget
    base = run_cpu_sample(a, max(a.steps, 1), min(max(a.warmup, 0), 1), budget_s=a.cpu_budget_s)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": base["steps_timed"], "steps_requested": a.steps, "warmup": min(max(a.warmup, 0), 1),
            "ms_per_step": base["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a),
                       "note": "CPU run on the N=%d instance of the workload's generator (a bounded sample: the 10 M-node graph takes "
                               "~5 min per step on the host); edges/s is size-normalised; timed steps are capped by --cpu-budget-s"
                               % min(a.cpu_sample_nodes, a.nodes)},
            "func_evals_per_sec": base["func_evals_per_sec"],
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# library baseline: the reference's arithmetic through ATen on the same B200
# --------------------------------------------------------------------------------------------------


def library_baseline(a, dev, blk, x, nnz, our_ms):
    """What the reference modules do once moved to the GPU with ``.cuda()``: ``F.group_norm`` + ``torch.cat`` + ``torch.mm``
    (cuBLAS SGEMM, TF32 off) + ``torch.spmm`` on the COO adjacency (cuSPARSE) + bias + relu (GCN/models.py:172-179,
    GCN/layers.py:69-75), and ``torch.autograd.grad`` over (y, theta) for one adjoint evaluation.  Plain ATen calls, timed
    with CUDA events after the product's own measurement; an rk4 fwd+bwd step is 5 forward + 4 adjoint evaluations, the
    stage axpys are not charged (favourable to the library)."""
    import torch
    import torch.nn.functional as F
    synth = _load_synth()
    torch.backends.cuda.matmul.allow_tf32 = False
    n, d = a.nodes, a.dim
    try:
        torch.cuda.empty_cache()
        row, col, val = synth.powerlaw_graph(n, avg_degree=a.avg_degree, locality=a.locality, seed=a.seed, device=dev)
        adj_raw = torch.sparse_coo_tensor(torch.stack([row, col]), val, (n, n))      # uncoalesced flag, as GCN/utils.py:222-229 builds it
        del row, col, val
        f = blk.odefunc
        W, b, gamma, beta = (f.gc1.weight.detach(), f.gc1.bias.detach(), f.norm1.weight.detach(), f.norm1.bias.detach())
        groups = f.norm1.num_groups

        def feval(adj, t, y, W, b, gamma, beta):
            xn = F.group_norm(y, groups, gamma, beta, 1e-5)
            ttx = torch.cat([torch.ones_like(xn[:, :1]) * t, xn], 1)
            return F.relu(torch.spmm(adj, torch.mm(ttx, W)) + b)

        def timed(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        y = x.detach()
        g = torch.randn_like(y)
        with torch.no_grad():
            ms_raw = timed(lambda: feval(adj_raw, 0.3, y, W, b, gamma, beta), reps=1)
        adj = adj_raw.coalesce()
        del adj_raw
        with torch.no_grad():
            ms_f = timed(lambda: feval(adj, 0.3, y, W, b, gamma, beta))

        def aug():
            ps = [p.detach().requires_grad_(True) for p in (W, b, gamma, beta)]
            yy = y.detach().requires_grad_(True)
            with torch.enable_grad():
                out = feval(adj, 0.3, yy, *ps)
                torch.autograd.grad(out, [yy] + ps, -g)

        ms_a = timed(aug)
        step_ms = 5 * ms_f + 4 * ms_a
        return {"kind": "reference arithmetic through ATen on the same B200 (cuSPARSE torch.spmm on the coalesced COO adjacency + cuBLAS "
                        "fp32 torch.mm + F.group_norm + torch.cat; torch.autograd.grad for the adjoint evaluation)",
                "fwd_func_eval_ms": ms_f, "adjoint_func_eval_ms": ms_a,
                "fwd_func_eval_ms_uncoalesced_as_the_reference_builds_adj": ms_raw,
                "step_ms_estimate": step_ms, "value": nnz * NFE_PER_STEP / (step_ms / 1e3), "unit": UNIT,
                "ours_over_library": step_ms / our_ms,
                "note": "5 forward + 4 adjoint evaluations per rk4 fwd+bwd step; Runge-Kutta axpys and the optimiser are not charged to the library"}
    except Exception as e:   # e.g. out of memory on a smaller GPU: the baseline is optional, the product's numbers are not
        return {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}


# --------------------------------------------------------------------------------------------------
# ours
# --------------------------------------------------------------------------------------------------


def run_ours(a):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        sys.path.insert(0, ROOT)
        import graph_odenet_b200  # noqa: F401
        from graph_odenet_b200 import parallel as _par
        _par.init_process_group(torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    import graph_odenet_b200  # noqa: F401  (fails loudly if libgode.so is missing)
    from graph_odenet_b200 import _lib, ops, synth
    from graph_odenet_b200.GCN import models

    n, d = a.nodes, a.dim
    row, col, val = synth.powerlaw_graph(n, avg_degree=a.avg_degree, locality=a.locality, seed=a.seed, device=dev)
    nnz = int(val.numel())                       # global stored entries of A_hat: the metric's "edges"
    relabel_info = None
    if a.shuffle_ids or a.reorder != "none":
        # pure relabellings of the same graph (values travel with their entries): every rank computes the same permutation
        from graph_odenet_b200 import parallel as _rl
        t_rl = time.perf_counter()
        before = _rl.halo_fraction(row, col, n, max(world, 8))
        if a.shuffle_ids:
            gsh = torch.Generator(device=dev).manual_seed(a.seed + 1)
            shuf = torch.randperm(n, generator=gsh, device=dev)
            row, col = shuf[row], shuf[col]
        shuffled = _rl.halo_fraction(row, col, n, max(world, 8))
        if a.reorder != "none":
            perm = _rl.locality_order(row, col, n, a.reorder, a.reorder_sweeps)
            row, col = _rl.relabel(row, col, perm)
            del perm
        torch.cuda.synchronize()
        relabel_info = {"halo_rows_per_owned_row_%dway" % max(world, 8): {"as_generated": before, "after_shuffle": shuffled,
                                                                          "after_reorder": _rl.halo_fraction(row, col, n, max(world, 8))},
                        "seconds": time.perf_counter() - t_rl}
    halo_info = None
    if world > 1:
        # contiguous row blocks of A_hat / A_hat^T, halo exchange of the gather operand per evaluation
        from graph_odenet_b200 import parallel
        plan = parallel.PartitionedPlan.build(row, col, val, n, rank, world, mode=a.halo_mode)
        lo, hi = plan.lo, plan.hi
        halo_info = {"halo_mode": plan.mode, "rows_owned": plan.n_rows, "halo_rows_fwd": plan.halo.n_halo, "halo_rows_bwd": plan.halo_t.n_halo,
                     "nvlink_bytes_per_step_rank0": plan.halo_bytes_per_step(d, 4, 4) if a.method == "rk4" else None}
    else:
        plan = ops.GraphPlan.from_coo(row, col, val, n, n)
        lo, hi = 0, n
    del row, col, val
    torch.cuda.empty_cache()
    torch.manual_seed(a.seed)
    blk = models.ODEBlock(models.ODEfunc(d), method=a.method).to(dev)
    opt = torch.optim.Adam(blk.parameters(), lr=0.01, weight_decay=5e-4)   # GCN/train_res.py:126-127
    gen = torch.Generator(device=dev).manual_seed(a.seed)
    x_dev = torch.randn(n, d, device=dev, generator=gen)
    if world > 1:
        x_dev = x_dev[lo:hi].clone()             # every rank draws the same [N, d] and keeps its rows
        torch.cuda.empty_cache()
    n_loc = hi - lo
    inv_count = 1.0 / (float(n) * d)

    def step(x):
        opt.zero_grad(set_to_none=True)
        xx = x.requires_grad_(True)
        y = blk(xx, plan)
        # loss = this rank's share of 0.5 * mean(y^2) over the whole graph, with its gradient y / (N d) formed directly
        # (one dot product + one scaled copy): the loss is not part of the hot path, and the autograd graph of
        # (y * y).sum() costs five elementwise passes over [N, d]
        with torch.no_grad():
            flat = y.reshape(-1)
            loss = 0.5 * inv_count * torch.dot(flat, flat)
            gy = y * inv_count
        y.backward(gy)                           # parameter gradients are summed over ranks inside the adjoint
        opt.step()
        return loss

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    clk = ClockSampler(local).start()
    blk.nfe = 0
    step(x_dev.detach())                         # first step: allocator growth, lazy kernel attributes
    sync()
    clk.mark_load()
    for _ in range(max(a.warmup, 3) - 1):
        step(x_dev.detach())
    sync()
    nfe_per_step = blk.nfe // max(a.warmup, 3)
    assert nfe_per_step == NFE_PER_STEP or a.method != "rk4", nfe_per_step

    # ---- timed region: device-resident inputs ---------------------------------------------------
    lib = _lib.lib
    lib.gode_profile_enable(1)
    launches0 = lib.gode_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    clk.mark_begin()
    ev0.record()
    for _ in range(a.steps):
        loss = step(x_dev.detach())
    ev1.record()
    sync()
    clk.mark_end()
    clk.stop()
    ms = ev0.elapsed_time(ev1) / a.steps
    if world > 1:
        plan.check_peers()                       # a timed-out device-side wait on a peer's flag invalidates the run

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms = max_over_ranks(ms)
    launches = int(lib.gode_launch_count() - launches0)
    import ctypes as C
    prof = {}
    for name, kind in (("agg_fwd", 0), ("agg_t", 1), ("transform", 2), ("vjp_dense", 3), ("other", 4)):
        cnt, tot, mx = C.c_int(), C.c_float(), C.c_float()
        lib.gode_profile_read(kind, C.byref(cnt), C.byref(tot), C.byref(mx))
        prof[name] = {"launches": cnt.value, "ms_total": tot.value, "ms_avg": tot.value / max(cnt.value, 1)}
    lib.gode_profile_enable(0)
    value = nnz * nfe_per_step / (ms / 1e3)

    # ---- roofline of the dominant kernel (the A_hat*S gather with fused epilogue) -----------------
    # Algorithmic bytes of one launch = CSR once + one [N,d] stream per operand the launch must read or write:
    # S, k_i (when a later stage needs it), y0, the k_j with a non-zero coefficient in the stage combination, y_next,
    # and in the adjoint the state a and gP.  Averaged over the launches of a step.  The compulsory-only
    # figure of SURVEY 8d (CSR + S + k, no Runge-Kutta operands) is reported next to it.
    peak, peak_src = peaks()
    from graph_odenet_b200 import odeint as _od
    n_rows_loc = n_loc
    n_gather = n_loc + (plan.halo.n_halo if world > 1 else 0)
    nnz_loc = plan.A.nnz if world > 1 else nnz
    csr_bytes = nnz_loc * 8 + (n_rows_loc + 1) * 4
    row_stream = n_rows_loc * d * 4
    b_compulsory = csr_bytes + n_gather * d * 4 + row_stream

    def launch_streams(method):
        """[N,d] streams (besides S) of every agg_fwd launch of one fwd+bwd step on grid [0,1]."""
        tab = _od.TABLEAUS[method]
        running = _od._running_final(tab)
        out = []
        ks = [None] * tab.s
        for aug in (False, True):
            if aug:
                out.append(1)                                        # f(t1): writes k only
            for i in range(tab.s):
                last = i == tab.s - 1
                _, kprev, _, _, sec = _od._stage_plan(tab, i, 1.0, ks, running)
                if aug and last:
                    out.append(2)                                    # the adjoint's last stage forms no y(t0): a, gP only
                    continue
                second = sec is not None and not aug                 # running final combination (the adjoint's y needs none)
                store = (not last) and _od._needed_later(tab, i) and sec is None
                # k_i, base state, k_j.., y_next, (running final combination), (a, gP)
                out.append((1 if store else 0) + 1 + len(kprev) + 1 + (1 if second else 0) + (2 if aug else 0))
        return out

    if a.method in _od.FIXED_METHODS:
        ls = launch_streams(a.method)
        b_f = sum(csr_bytes + n_gather * d * 4 + k * row_stream for k in ls) / len(ls)
    else:
        b_f = b_compulsory
    agg = prof["agg_fwd"]
    sec = agg["ms_avg"] / 1e3
    # achieved = the COMPULSORY bytes of one function evaluation (CSR once + S once + k once) / launch time
    achieved = b_compulsory / sec / 1e9 if agg["launches"] else None
    traffic = l2cap = None
    if world == 1 and n == 10_000_000 and d == 128:
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("agg_fwd_dram_bytes_per_launch")
        except Exception:
            traffic = None
    try:    # measured L2 -> SM delivery rate of random 512 B rows on this pool's B200 (tools/l2_gather_bench.cu)
        l2cap = json.load(open(os.path.join(ROOT, "profiles", "l2_gather.json")))["band_TBps"]
    except Exception:
        l2cap = None
    l2_bytes = nnz_loc * d * 4.0             # every stored entry brings one d-float neighbour row from L2 to an SM
    roofline = {"bound": "hbm", "kernel": "k_spmm_t2<32,8> (A_hat*S gather + bias + relu + RK combine)%s" % (
                    "" if world == 1 else " on rank 0's row block"),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic,
                "traffic_note": None if traffic is None else (
                    "ncu dram__bytes_read + dram__bytes_write, average of five consecutive launches of one step (profiles/traffic.json): "
                    "25.4 GB for the bare gather incl. hub rows (2.15x the compulsory bytes: the non-local 10 % of the entries and "
                    "the hubs' bands miss L2) + 5.1 GB per [N,d] operand of the fused Runge-Kutta / adjoint epilogue, which the "
                    "compulsory count of SURVEY 8d leaves out (with_fused_epilogue_operands counts them)"),
                "algorithmic_bytes_per_launch": b_compulsory, "ms_per_launch": agg["ms_avg"],
                "launches_timed": agg["launches"], "peak_source": peak_src,
                "model": "SURVEY 8d compulsory bytes per function evaluation: nnz*8 + (N+1)*4 + 2*N*d*4",
                "with_fused_epilogue_operands": {
                    "bytes_per_launch": b_f, "frac": (b_f / sec / 1e9 / peak) if agg["launches"] else None,
                    "note": "secondary: also counts the [N,d] Runge-Kutta / adjoint operands the fused epilogue streams "
                            "(y0, k_j, y_next, a, gP), averaged over the launches of a step"},
                "l2_to_sm": {"bytes_per_launch": l2_bytes, "achieved_TBps": (l2_bytes / sec / 1e12) if agg["launches"] else None,
                             "measured_ceiling_TBps": l2cap,
                             "frac": (l2_bytes / sec / 1e12 / l2cap) if (agg["launches"] and l2cap) else None,
                             "note": "what actually bounds the gather: nnz*d*4 bytes of neighbour rows cross the L2 -> SM fabric per "
                                     "launch whatever the L2 hit rate; ceiling = tools/l2_gather_bench.cu on this pool (profiles/l2_gather.json)"},
                "whole_step": {"algorithmic_bytes": 5 * b_compulsory + 4 * (2 * csr_bytes + 6 * row_stream) if a.method == "rk4" else None,
                               "frac": ((5 * b_compulsory + 4 * (2 * csr_bytes + 6 * row_stream)) / (ms / 1e3) / 1e9 / peak)
                               if a.method == "rk4" else None,
                               "note": "SURVEY 8d: 5*B_f + 4*B_aug per rk4 fwd+bwd over the whole step time"},
                "step_share": agg["ms_total"] / (ms * a.steps), "kernel_classes_ms": prof}

    # ---- e2e: public API, host buffers ------------------------------------------------------------
    e2e = None
    if not a.no_e2e:
        # Input pipeline of a training loop: the step's input lives in pinned host memory and is copied to the device every
        # step; the copy of step k+1 is issued on a copy stream while step k computes (two staging buffers), exactly what a
        # DataLoader with pin_memory + non_blocking prefetch does.  Every step still pays its own 5.12 GB host->device copy
        # and a device->host read of its loss inside the timed region.
        x_host = torch.empty(n_loc, d, dtype=torch.float32, pin_memory=True)
        x_host.copy_(x_dev)
        x_stage = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
        copy_stream = torch.cuda.Stream(device=dev)

        def issue_copy(k):
            with torch.cuda.stream(copy_stream):
                x_stage[k & 1].copy_(x_host, non_blocking=True)

        def e2e_step(k):
            torch.cuda.current_stream().wait_stream(copy_stream)      # input k has arrived
            issue_copy(k + 1)                                         # its buffer was last read by step k-1, which has finished
            return float(step(x_stage[k & 1].detach()).item())

        issue_copy(0)
        e2e_step(0)                                                   # warm-up; leaves copy 1 in flight
        sync()
        t0 = time.perf_counter()
        for k in range(1, a.steps + 1):                               # steady state: one copy issued and one awaited per step
            e2e_step(k)
        sync()                                                        # includes the last copy issued
        e_ms = max_over_ranks((time.perf_counter() - t0) / a.steps * 1e3)
        e2e = {"value": nnz * nfe_per_step / (e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": n * d * 4,
               "d2h_bytes_per_step": 4 * world, "ms_per_step": e_ms,
               "note": "module API (ODEBlock forward / backward + Adam) on host buffers: each step's input is copied from "
                       "pinned host memory (the copy of step k+1 overlaps step k, two staging buffers) and its loss is read "
                       "back; graph plan stays resident across steps as adj.cuda() does in GCN/train_res.py:57; bytes are "
                       "summed over ranks (each rank copies its own rows)"}
        del x_host, x_stage

    lib_base = None
    if not a.no_library_baseline and world == 1:
        lib_base = library_baseline(a, dev, blk, x_dev, nnz, ms)

    cpu = None
    if not a.no_cpu_baseline and world == 1:
        cpu = run_cpu_sample(a, steps=1, warmup=0)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    loss_total = loss.detach().clone()
    if world > 1:
        dist.all_reduce(loss_total)
    if rank != 0:
        dist.destroy_process_group()
        return
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "nnz": nnz, "solver": a.method, "func_evals_per_step": nfe_per_step,
                       "l2": "inputs larger than L2 (every [N,d] tensor is %.2f GB)" % (n * d * 4 / 1e9),
                       "optimizer": "Adam on the ODE function's parameters (in the timed region)",
                       "relabelling": relabel_info,
                       "partition": None if world == 1 else dict(
                           halo_info,
                           scheme="contiguous row blocks; " + (
                               ("the gathers that produce a stage state / a masked adjoint store the rows their peers reference "
                                "into the peers' halo tails over NVLink peer memory from their epilogues; every rank transforms "
                                "[owned | halo] rows itself (no exchange of the support between two gathers)"
                                if os.environ.get("GODE_PUSH_Y", "1") != "0" and plan.mode == "p2p-fused" else
                                "halo exchange of the gather operand per evaluation (libgode kernels storing rows into the peers' "
                                "halo tails over NVLink peer memory)")
                               if plan.mode.startswith("p2p") else
                               "halo exchange of the gather operand per evaluation (pack + NCCL all-to-all-v)"),
                           # what the exchange leaves exposed on rank 0: the step minus the CUDA-event time of the libgode compute
                           # classes (gathers, transforms, dense VJP chain) -- flag waits, pushes not hidden, loss, Adam
                           ms_outside_compute_classes=ms - sum(v["ms_total"] for v in prof.values()) / a.steps)},
            "func_evals_per_sec": nfe_per_step / (ms / 1e3), "final_loss": float(loss_total.item()),
            "roofline": roofline, "cpu_baseline": cpu, "library_baseline": lib_base, "e2e": e2e, "gpu_launches": launches,
            "clocks": clk.summary()}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    return run_ours(a)


if __name__ == "__main__":
    main()
