#!/usr/bin/env python
"""Timings of BASELINE.json's other configurations on ONE B200 -- the parity-test cases, measured so that
DESIGN.md can quote them; the headline metric stays bench.py (config 4).  One JSON line per case on stdout.

  config 1  GCN-ODE (ODEGCN3, dopri5 tol 1e-5 as GCN/train_res.py) on Cora, per-epoch time + nfe; the reference arithmetic
            on the host CPU (oracle port) beside it
  config 2  Pubmed: discrete residual (res3) vs ODE (ode3) with rk4 and dopri5, hidden 16 / 64 / 128
  config 3  GAT-ODE block, 8 heads x 16, on a Citeseer-shaped synthetic graph scaled to 1 M nodes (1 405 470 edges)
  config 5  QC edge-conditioned model (EdgeGCN_K_Sum, K = 3, hidden 73) on QM9-shaped synthetic molecule batches

Small graphs are launch / latency bound (state of a few MB, L2 resident): time per epoch and per function evaluation is
the figure of merit there, not a roofline fraction.
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import graph_odenet_b200  # noqa: E402,F401
from graph_odenet_b200 import synth, train, utils  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def emit(**kw):
    print(json.dumps(kw), flush=True)


def epochs(family, argv, n_warm=3):
    """Runs the reference's training loop (graph-odenet_b200/train.py) and returns per-epoch wall times after warm-up."""
    lines = []
    res = train.main(family, argv, out=lambda *a, **k: lines.append(" ".join(str(x) for x in a)))
    t = [float(l.split("time: ")[1].split("s")[0]) for l in lines if l.startswith("Epoch")]
    nfe = [(h[1], h[2]) for h in res["history"]]
    t = t[n_warm:]
    return {"epoch_ms_median": 1e3 * sorted(t)[len(t) // 2], "epoch_ms_min": 1e3 * min(t), "epochs_timed": len(t),
            "nfe_f_last": nfe[-1][0], "nfe_b_last": nfe[-1][1], "test_acc": res["acc"]}


def config1():
    npz = os.path.join(GOLD, "planetoid_cora.npz")
    for method, extra in (("dopri5", []), ("rk4", [])):
        r = epochs("GCN", ["--model", "ode3", "--dataset", "cora", "--npz", npz, "--epochs", "30", "--method", method] + extra)
        emit(config=1, case="ODEGCN3 on Cora, %s, 30 epochs (train + eval forward per epoch, as GCN/train_res.py)" % method,
             device="B200", **r)
    # the reference arithmetic on the host (oracle port: reference modules' formulas + restated torchdiffeq)
    from oracle import gcn_ref
    data = utils.load_npz(npz, "GCN")
    adj, feats, labels, idx_train = data[0], data[1], data[2], data[3]
    torch.manual_seed(42)
    from graph_odenet_b200.GCN import models
    m = models.ODEGCN3(feats.shape[1], 16, int(labels.max()) + 1, 0.5)
    p = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    torch.set_num_threads(os.cpu_count() or 1)
    times, nfe = [], None
    for _ in range(4):
        t0 = time.perf_counter()
        st = {}
        out, _ = gcn_ref.odegcn3(feats, adj, p, tol=1e-5, stats=st)
        loss = torch.nn.functional.nll_loss(out[idx_train], labels[idx_train])
        loss.backward()
        times.append(time.perf_counter() - t0)
        nfe = st
    emit(config=1, case="same model, reference arithmetic on the host CPU (oracle port), fwd+bwd only, dropout off",
         device="cpu x%d" % (os.cpu_count() or 1), step_ms_median=1e3 * sorted(times[1:])[len(times[1:]) // 2],
         solver_stats={k: (v if isinstance(v, (int, float)) else str(v)) for k, v in (nfe or {}).items()})


def _pubmed_npz():
    """The reference's data/ has the Pubmed graph but not its feature file: the fixture holds the real
    A_hat; features are synthetic TF-IDF-like rows (about 50 non-zeros of 500, row-normalised as GCN/utils.py:185), labels
    synthetic (3 classes), the Planetoid split sizes (60 / 500 / 1000).  Written once to a temporary .npz for train.main."""
    import tempfile
    import numpy as np
    import scipy.sparse as sp
    c = dict(np.load(os.path.join(GOLD, "planetoid_pubmed.npz")))
    n, f = int(c["n"]), 500
    rng = np.random.default_rng(0)
    cols = rng.integers(0, f, size=(n, 50))
    rows = np.repeat(np.arange(n), 50)
    x = sp.csr_matrix((np.ones(n * 50, dtype=np.float32), (rows, cols.ravel())), shape=(n, f))
    x.sum_duplicates()
    x.data[:] = 1.0
    x = sp.diags(1.0 / np.asarray(x.sum(1)).ravel()).dot(x).tocsr().astype(np.float32)
    c.update(feat_data=x.data, feat_indices=x.indices, feat_indptr=x.indptr, nfeat=f,
             labels=rng.integers(0, 3, size=n), idx_train=np.arange(60), idx_val=np.arange(60, 560),
             idx_test=np.arange(n - 1000, n))
    path = os.path.join(tempfile.gettempdir(), "gode_pubmed_synth.npz")
    np.savez(path, **c)
    return path


def config2():
    npz = _pubmed_npz()
    for hidden in (16, 64, 128):
        for model, method in (("res3", None), ("ode3", "rk4"), ("ode3", "dopri5")):
            argv = ["--model", model, "--dataset", "pubmed", "--npz", npz, "--epochs", "20", "--hidden", str(hidden)]
            if method:
                argv += ["--method", method]
            r = epochs("GCN", argv)
            emit(config=2, case="%s%s on Pubmed (real graph, synthetic features / labels), hidden %d" % (
                model, "/" + method if method else "", hidden), device="B200", **r)


def config3():
    from graph_odenet_b200.GAT import models
    dev = torch.device("cuda:0")
    n = 1_000_000
    # Citeseer-shaped: E = round(4676 / 3327 * N) directed edges, one per undirected pair (src < tgt), power-law degrees
    row, col = synth.powerlaw_graph(n, avg_degree=1 + 2 * 4676 / 3327, seed=0, device=dev, return_raw=True)
    keep = row < col
    src, tgt = row[keep].contiguous(), col[keep].contiguous()
    d, heads = 128, 8
    torch.manual_seed(0)
    blk = models.ODEBlock(models.ODEfunc(d, heads=heads), method="rk4").to(dev)
    x = torch.randn(n, d, device=dev)
    g = torch.randn(n, d, device=dev) / n

    def step():
        for p in blk.parameters():
            p.grad = None
        xx = x.clone().requires_grad_(True)
        blk.nfe = 0
        y = blk(xx, src, tgt, None)
        y.backward(g)
        return blk.nfe

    for _ in range(3):
        nfe = step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    k = 5
    for _ in range(k):
        step()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / k
    e = int(src.numel())
    emit(config=3, case="GAT-ODE block (8 heads x 16), rk4 fwd+bwd, Citeseer-shaped synthetic graph", device="B200",
         nodes=n, edges=e, d=d, heads=heads, ms_per_step=ms, func_evals_per_step=nfe, func_evals_per_sec=nfe / (ms / 1e3),
         edges_per_sec=e * nfe / (ms / 1e3))


def config5():
    from graph_odenet_b200.QC import layer_models
    dev = torch.device("cuda:0")
    for n_mol in (20, 4096):
        b = synth.qm9_like_batch(n_mol, 73, seed=0, device=dev)
        torch.manual_seed(0)
        model = layer_models.EdgeGCN_K_Sum(node_features=13, edge_features=5, target_features=12, hidden_features=73,
                                           num_layers=3).to(dev)
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
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        k = 10
        for _ in range(k):
            step()
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / k
        emit(config=5, case="EdgeGCN_K_Sum (K = 3, hidden 73) training step on QM9-shaped synthetic molecules", device="B200",
             molecules=n_mol, atoms=int(b["node_features"].shape[0]), edges=int(b["esrc"].numel()), ms_per_step=ms,
             molecules_per_sec=n_mol / (ms / 1e3))


def config5dp(n_mol_per_rank=8192):
    """BASELINE config 5 as it shards: molecules split data-parallel over the ranks of a torchrun launch, full
    model replica per rank, one gradient all-reduce (57 MB) per step.  Weak scaling: every rank gets n_mol_per_rank molecules.

        python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_configs.py 5dp"""
    import torch.distributed as dist
    from graph_odenet_b200 import parallel
    from graph_odenet_b200.QC import layer_models
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        parallel.init_process_group(dev)
    b = synth.qm9_like_batch(n_mol_per_rank, 73, seed=rank, device=dev)
    torch.manual_seed(0)                                                   # identical replicas
    model = layer_models.EdgeGCN_K_Sum(node_features=13, edge_features=5, target_features=12, hidden_features=73, num_layers=3).to(dev)
    target = torch.randn(n_mol_per_rank, 12, device=dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    t_ar = [0.0]

    def step():
        opt.zero_grad(set_to_none=True)
        out = model(b["node_features"], b["edge_features"], b["esrc"], b["etgt"], b["batch"], batch_size=n_mol_per_rank)
        torch.nn.functional.mse_loss(out, target).backward()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        parallel.allreduce_gradients(model.parameters(), local_weight=1.0 / world)
        e1.record()
        opt.step()
        return e0, e1

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    k, ars = 5, []
    for _ in range(k):
        ars.append(step())
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1) / k, sum(a.elapsed_time(b_) for a, b_ in ars) / k], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        emit(config=5, case="EdgeGCN_K_Sum (K = 3, hidden 73) data-parallel training step, gradient all-reduce over NCCL", device="B200",
             n_gpus=world, molecules_per_rank=n_mol_per_rank, molecules=n_mol_per_rank * world, ms_per_step=float(ms[0]),
             allreduce_ms=float(ms[1]), grad_bytes=4 * sum(p.numel() for p in model.parameters()),
             molecules_per_sec=n_mol_per_rank * world / (float(ms[0]) / 1e3))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "2", "3", "5"]
    for w in which:
        try:
            {"1": config1, "2": config2, "3": config3, "5": config5, "5dp": config5dp}[w]()
        except Exception as e:  # noqa: BLE001 -- a failing case must not hide the others
            emit(config=str(w), error="%s: %s" % (type(e).__name__, e))
