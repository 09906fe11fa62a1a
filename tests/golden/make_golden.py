#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference modules are imported from where they lie (``/root/reference/{GCN,GAT,QC}``); nothing is
copied.  Two shims are needed to import them under this image and both are import-only:

* ``torchdiffeq`` is absent -> ``sys.modules['torchdiffeq']`` is pointed at ``oracle/odeint.py`` (the
  restated solver; parity for the solver itself is therefore *unpinned*, see oracle/odeint.py);
* ``GCN/utils.py`` imports matplotlib and a removed scipy path at module level (utils.py:5-8) ->
  empty stub modules satisfy the import; the loader body (utils.py:154-202) then runs unmodified.

Outputs (all small, committed):
  planetoid_<ds>.npz   reference loader output (adjacency COO, GAT edge list, labels, splits, features)
  gcn_golden.npz       GraphConvolution / ODEfunc / ODEfunc2 / ODEBlock / whole-model outputs + grads
  gat_golden.npz       GAT GraphConvolution + ODEfunc outputs + grads
  qc_golden.npz        QC EdgeGraphConvolution / EdgeEncoderMLP / EdgeGCN_K_Sum outputs + grads
  qc_collate_golden.npz  the DataLoader collate of QC/datasets/utils.py on synthetic molecules (``--only-qc-collate``)
  set2set_golden.npz   QC Set2Set readout and EdgeGCN_K_Set2Set outputs + grads (``--only-set2set`` regenerates it alone)
"""
from __future__ import annotations

import functools
import importlib
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import odeint as restated  # noqa: E402
from oracle import graph_ops  # noqa: E402

STATS = {}


def _install_shims():
    td = types.ModuleType("torchdiffeq")
    td.odeint = restated.odeint
    td.odeint_adjoint = restated.odeint_adjoint
    sys.modules["torchdiffeq"] = td
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors",
                 "scipy.sparse.linalg.eigen", "scipy.sparse.linalg.eigen.arpack"]:
        m = types.ModuleType(name)
        sys.modules[name] = m
    sys.modules["matplotlib.colors"].colorConverter = None
    sys.modules["scipy.sparse.linalg.eigen.arpack"].eigsh = None


def load_ref(subdir, names=("layers", "models", "utils")):
    """Import the reference's bare-named modules from one experiment directory."""
    for n in ("layers", "models", "utils", "mpnn", "set2set", "torch_scatter", "layer_models"):
        sys.modules.pop(n, None)
    sys.path.insert(0, os.path.join(REF, subdir))
    try:
        mods = types.SimpleNamespace(**{n: importlib.import_module(n) for n in names})
    finally:
        sys.path.pop(0)
    for n in ("layers", "models", "utils", "mpnn", "set2set", "torch_scatter", "layer_models"):
        sys.modules.pop(n, None)
    return mods


def rnd(seed, *shape, scale=1.0):
    """Inputs come from numpy's legacy RandomState (stream frozen by numpy policy) so the fixtures only
    need to hold the seed; tests regenerate the same float32 arrays with ``tests/_golden.py:rnd``."""
    return torch.from_numpy((np.random.RandomState(seed).standard_normal(shape) * scale).astype(np.float32))


def sd_np(module, prefix=""):
    return {prefix + k: v.detach().numpy().copy() for k, v in module.state_dict().items()}


def set_method(models_mod, method, options=None, stats=None):
    """Route the reference ODEBlock's odeint call (models.py:192) to a given solver method."""
    models_mod.odeint = functools.partial(restated.odeint_adjoint, method=method, options=options, stats=stats)


def main():
    _install_shims()
    gcn = load_ref("GCN")
    gat = load_ref("GAT")
    cwd = os.getcwd()
    os.chdir(REF)

    # ------------------------------------------------------------------ planetoid loader output
    import pickle as pkl
    import scipy.sparse as sp
    data = {}
    for ds in ["cora", "citeseer", "pubmed"]:
        out = {}
        with open("data/ind.%s.graph" % ds, "rb") as f:
            graph = pkl.load(f, encoding="latin1")
        # reference pipeline, verbatim calls (GCN/utils.py:180,186,222-229)
        import networkx as nx
        adj_raw = nx.adjacency_matrix(nx.from_dict_of_lists(graph))
        adj = gcn.utils.normalize(adj_raw + sp.eye(adj_raw.shape[0]))
        adj_t = gcn.utils.sparse_mx_to_torch_sparse_tensor(adj)
        idx, val = adj_t._indices().numpy(), adj_t._values().numpy()
        out.update(n=np.int64(adj_raw.shape[0]), coo_row=idx[0].astype(np.int32), coo_col=idx[1].astype(np.int32),
                   coo_val=val.astype(np.float32))
        raw = adj_raw.tocoo()
        out.update(raw_row=raw.row.astype(np.int32), raw_col=raw.col.astype(np.int32),
                   raw_val=raw.data.astype(np.float32))
        # oracle restatement of the same pipeline must agree bit for bit
        n, r, c = graph_ops.adjacency_from_dict_of_lists(graph)
        assert n == adj_raw.shape[0]
        rr, cc, vv = graph_ops.add_self_loops(n, r, c, np.ones(len(r)))
        vv = graph_ops.normalize_rows(n, rr, cc, vv)
        rr, cc, vv = graph_ops.to_coo_f32(rr, cc, vv)
        o1 = np.lexsort((idx[1], idx[0]))
        assert np.array_equal(rr, idx[0][o1]) and np.array_equal(cc, idx[1][o1]), ds
        assert np.array_equal(vv.view(np.uint32), val[o1].view(np.uint32)), ds
        # GAT edge list (GAT/utils.py:187-196)
        G = nx.from_dict_of_lists(graph)
        edges = np.array(G.edges, dtype=np.int64).reshape(-1, 2)
        out.update(gat_src=edges[:, 0].astype(np.int32), gat_tgt=edges[:, 1].astype(np.int32))
        if ds != "pubmed":  # ind.pubmed.allx is missing from the reference checkout
            a, feats, labels, itr, iva, ite = gcn.utils.load_data_new(ds)
            assert torch.equal(a._indices(), adj_t._indices()) and torch.equal(a._values(), adj_t._values())
            fs = sp.csr_matrix(feats.numpy())
            out.update(feat_indptr=fs.indptr.astype(np.int32), feat_indices=fs.indices.astype(np.int32),
                       feat_data=fs.data.astype(np.float32), nfeat=np.int64(feats.shape[1]),
                       labels=labels.numpy().astype(np.int16), idx_train=itr.numpy().astype(np.int32),
                       idx_val=iva.numpy().astype(np.int32), idx_test=ite.numpy().astype(np.int32))
            data[ds] = (a, feats, labels, itr)
        else:
            data[ds] = (adj_t, None, None, None)
        np.savez_compressed(os.path.join(HERE, "planetoid_%s.npz" % ds), **out)
        print(ds, "N", out["n"], "nnz", len(val), "gatE", len(edges))
    os.chdir(cwd)

    adj, feats, labels, idx_train = data["cora"]
    N = adj.shape[0]
    G = {}
    # a 512-node induced subgraph of Cora, normalised by the reference's own functions: keeps d=128 fixtures small
    c = np.load(os.path.join(HERE, "planetoid_cora.npz"))
    raw = sp.coo_matrix((c["raw_val"], (c["raw_row"], c["raw_col"])), shape=(N, N)).tocsr()[:512, :512]
    adj_s = gcn.utils.sparse_mx_to_torch_sparse_tensor(gcn.utils.normalize(raw + sp.eye(512)))
    G.update({"sub/row": adj_s._indices()[0].numpy().astype(np.int32), "sub/col": adj_s._indices()[1].numpy().astype(np.int32),
              "sub/val": adj_s._values().numpy()})
    NS = 512

    # ------------------------------------------------------------------ GraphConvolution fwd + grads
    # (inputs are regenerated in the tests from the recorded seeds: rnd(seed, shape))
    torch.manual_seed(42)
    layer = gcn.layers.GraphConvolution(32, 16)
    x = rnd(1, N, 32).requires_grad_(True)
    g = rnd(2, N, 16)
    y = layer(x, adj)
    y.backward(g)
    G.update({"gc/out": y.detach().numpy(), "gc/grad_x": x.grad.numpy(),
              "gc/grad_weight": layer.weight.grad.numpy(), "gc/grad_bias": layer.bias.grad.numpy(),
              **sd_np(layer, "gc/p/")})

    # ------------------------------------------------------------------ ODEfunc / ODEfunc2 fwd + VJP
    for d, A_, n_ in ((16, adj, N), (128, adj_s, NS)):
        torch.manual_seed(100 + d)
        f = gcn.models.ODEfunc(d)
        with torch.no_grad():  # non-trivial affine so gamma/beta grads are exercised
            f.norm1.weight.uniform_(0.5, 1.5)
            f.norm1.bias.uniform_(-0.5, 0.5)
        f.set_adj(A_)
        x = rnd(10 + d, n_, d).requires_grad_(True)
        t = torch.tensor(0.37, requires_grad=True)
        g = rnd(20 + d, n_, d)
        y = f(t, x)
        grads = torch.autograd.grad(y, (x, t) + tuple(f.parameters()), g)
        k = "odefunc%d/" % d
        G.update({k + "out": y.detach().numpy(), k + "grad_x": grads[0].numpy(), k + "grad_t": grads[1].numpy(),
                  **sd_np(f, k + "p/")})
        for (name, _), gr in zip(f.named_parameters(), grads[2:]):
            G[k + "grad/" + name] = gr.numpy()

    torch.manual_seed(7)
    f2 = gcn.models.ODEfunc2(128, 0.0)
    f2.set_adj(adj_s)
    x = rnd(31, NS, 128).requires_grad_(True)
    t = torch.tensor(0.61, requires_grad=True)
    g = rnd(32, NS, 128)
    y = f2(t, x)
    grads = torch.autograd.grad(y, (x, t) + tuple(f2.parameters()), g)
    G.update({"odefunc2/out": y.detach().numpy(), "odefunc2/grad_x": grads[0].numpy(),
              "odefunc2/grad_t": grads[1].numpy(), **sd_np(f2, "odefunc2/p/")})
    for (name, _), gr in zip(f2.named_parameters(), grads[2:]):
        G["odefunc2/grad/" + name] = gr.numpy()

    # ------------------------------------------------------------------ ODEBlock: fixed-step + dopri5, fwd + adjoint grads
    cases = [(16, "cora", "rk4", None), (16, "cora", "dopri5", None),
             (16, "sub", "rk4", {"step_size": 0.25}), (16, "sub", "euler", {"step_size": 0.5}),
             (16, "sub", "midpoint", None), (128, "sub", "rk4", None), (128, "sub", "dopri5", None),
             (64, "sub", "rk4", None), (64, "sub", "dopri5", None)]
    for d, gname, method, opts in cases:
        A_, n_ = (adj, N) if gname == "cora" else (adj_s, NS)
        tag = method + ("" if not opts else "_h%g" % opts["step_size"])
        torch.manual_seed(200 + d)
        blk = gcn.models.ODEBlock(gcn.models.ODEfunc(d))
        with torch.no_grad():
            blk.odefunc.norm1.weight.uniform_(0.5, 1.5)
            blk.odefunc.norm1.bias.uniform_(-0.5, 0.5)
        x = rnd(40 + d, n_, d, scale=0.5).requires_grad_(True)
        g = rnd(50 + d, n_, d, scale=1.0 / n_)
        stats = {}
        set_method(gcn.models, method, opts, stats)
        blk.nfe = 0
        y = blk(x, A_)
        nfe_f = blk.nfe
        blk.nfe = 0
        y.backward(g)
        nfe_b = blk.nfe
        k = "odeblock%d_%s_%s/" % (d, gname, tag)
        G.update({k + "out": y.detach().numpy(), k + "grad_x": x.grad.numpy(),
                  k + "nfe_f": np.int64(nfe_f), k + "nfe_b": np.int64(nfe_b),
                  k + "acc_f": np.int64(stats["forward"].get("accepted", 0)),
                  k + "rej_f": np.int64(stats["forward"].get("rejected", 0)),
                  k + "acc_b": np.int64(stats["backward"].get("accepted", 0)),
                  k + "rej_b": np.int64(stats["backward"].get("rejected", 0)),
                  **sd_np(blk, k + "p/")})
        for name, p in blk.named_parameters():
            G[k + "grad/" + name] = p.grad.numpy().copy()
        print(k, "nfe", nfe_f, nfe_b, stats)

    # ------------------------------------------------------------------ whole models on Cora (eval mode)
    for name, cls, method in (("GCN3", gcn.models.GCN3, None), ("RGCN3", gcn.models.RGCN3, None),
                              ("RGCN3norm", gcn.models.RGCN3norm, None),
                              ("ODEGCN3_dopri5", gcn.models.ODEGCN3, "dopri5"),
                              ("ODEGCN3_rk4", gcn.models.ODEGCN3, "rk4")):
        torch.manual_seed(42)
        if method:
            set_method(gcn.models, method, None, None)
        model = cls(nfeat=feats.shape[1], nhid=16, nclass=int(labels.max()) + 1, dropout=0.5)
        model.eval()
        if method:
            model.nfe = 0
        out = model(feats, adj)
        nfe_f = model.nfe if method else 0
        if method:
            model.nfe = 0
        loss = torch.nn.functional.nll_loss(out[idx_train], labels[idx_train])
        loss.backward()
        k = "model_%s/" % name
        G.update({k + "out": out.detach().numpy(), k + "loss": np.float32(loss.item()),
                  k + "nfe_f": np.int64(nfe_f), k + "nfe_b": np.int64(model.nfe if method else 0),
                  **sd_np(model, k + "p/")})
        for pn, p in model.named_parameters():
            G[k + "grad/" + pn] = p.grad.numpy().copy()
        print(k, "loss", loss.item(), "nfe", nfe_f, model.nfe if method else 0)
    np.savez_compressed(os.path.join(HERE, "gcn_golden.npz"), **G)

    # ------------------------------------------------------------------ GAT (Cora edge list)
    src = torch.from_numpy(c["gat_src"].astype(np.int64))
    tgt = torch.from_numpy(c["gat_tgt"].astype(np.int64))
    E = len(src)
    Mtgt = torch.sparse_coo_tensor(torch.stack([tgt, torch.arange(E)]), torch.ones(E), (N, E))
    A = {}
    torch.manual_seed(11)
    layer = gat.layers.GraphConvolution(16, 8)
    x = rnd(61, N, 16).requires_grad_(True)
    g = rnd(62, N, 8)
    y = layer(x, src, tgt, Mtgt)
    y.backward(g)
    A.update({"gc/out": y.detach().numpy(), "gc/grad_x": x.grad.numpy(), **sd_np(layer, "gc/p/")})
    for pn, p in layer.named_parameters():
        A["gc/grad/" + pn] = p.grad.numpy().copy()
    torch.manual_seed(12)
    f = gat.models.ODEfunc(16)
    f.set_adj(src, tgt, Mtgt)
    x = rnd(63, N, 16).requires_grad_(True)
    t = torch.tensor(0.25, requires_grad=True)
    g = rnd(64, N, 16)
    y = f(t, x)
    grads = torch.autograd.grad(y, (x, t) + tuple(f.parameters()), g)
    A.update({"odefunc/out": y.detach().numpy(), "odefunc/grad_x": grads[0].numpy(),
              "odefunc/grad_t": grads[1].numpy(), **sd_np(f, "odefunc/p/")})
    for (pn, _), gr in zip(f.named_parameters(), grads[2:]):
        A["odefunc/grad/" + pn] = gr.numpy()
    np.savez_compressed(os.path.join(HERE, "gat_golden.npz"), **A)

    # ------------------------------------------------------------------ QC
    qc = load_ref("QC", names=("layers",))
    Q = {}
    for nf, nN, nE in ((24, 60, 130), (73, 18, 16)):
        torch.manual_seed(300 + nf)
        layer = qc.layers.EdgeGraphConvolution(nf, nf)
        x = rnd(70 + nf, nN, nf).requires_grad_(True)
        rs = np.random.RandomState(80 + nf)
        esrc = torch.from_numpy(rs.randint(0, nN, nE).astype(np.int64))
        etgt = torch.from_numpy(rs.randint(0, nN, nE).astype(np.int64))
        Etgt = torch.zeros(nN, nE)
        Etgt[etgt, torch.arange(nE)] = 1.0
        ed = rnd(90 + nf, nE, nf, nf, scale=1.0 / nf ** 0.5).requires_grad_(True)
        g = rnd(95 + nf, nN, nf)
        y = layer(x, esrc, Etgt, ed)
        y.backward(g)
        k = "egc%d/" % nf
        Q.update({k + "esrc": esrc.numpy().astype(np.int32), k + "etgt": etgt.numpy().astype(np.int32),
                  k + "out": y.detach().numpy(), k + "grad_x": x.grad.numpy(),
                  k + "grad_edge_data": ed.grad.numpy(), k + "grad_weight": layer.weight.grad.numpy(),
                  k + "grad_bias": layer.bias.grad.numpy(), **sd_np(layer, k + "p/")})
    torch.manual_seed(5)
    ee = qc.layers.EdgeEncoderMLP(5, 8)
    e = rnd(99, 20, 5)
    Q.update({"ee/out": ee(e).detach().numpy(), **sd_np(ee, "ee/p/")})
    np.savez_compressed(os.path.join(HERE, "qc_golden.npz"), **Q)
    make_set2set()
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(HERE, fn)) // 1024, "KiB")


def make_set2set():
    """set2set_golden.npz: the reference's Set2Set readout (QC/set2set.py:9-77) and EdgeGCN_K_Set2Set
    (QC/layer_models.py:125-163) on seeded inputs -- outputs and gradients."""
    _install_shims()
    qc = load_ref("QC", names=("set2set", "layer_models"))
    S = {}
    torch.manual_seed(11)
    C_, steps = 24, 3
    s2s = qc.set2set.Set2Set(C_, steps, num_layers=1)
    sizes = [7, 1, 12, 3, 18, 9]                       # graphs of very different sizes, one singleton
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    x = rnd(401, int(batch.numel()), C_).requires_grad_(True)
    g = rnd(402, len(sizes), 2 * C_)
    out = s2s(x, batch)
    out.backward(g)
    S.update({"s2s/x": x.detach().numpy(), "s2s/batch": batch.numpy().astype(np.int32), "s2s/g": g.numpy(),
              "s2s/out": out.detach().numpy(), "s2s/grad_x": x.grad.numpy(), **sd_np(s2s, "s2s/p/"),
              **{"s2s/grad/" + k: v.grad.numpy() for k, v in s2s.named_parameters()}})
    # whole model, hidden 24, K = 3, regression head with 12 targets
    torch.manual_seed(12)
    model = qc.layer_models.EdgeGCN_K_Set2Set(node_features=13, edge_features=5, target_features=12, hidden_features=24,
                                              num_layers=3, s2s_processing_steps=3, type="regression", dropout=0.0)
    model.eval()
    nN = int(batch.numel())
    rs = np.random.RandomState(13)
    # edges inside each molecule (block diagonal), a few per node
    offs = np.concatenate([[0], np.cumsum(sizes)])
    es, et = [], []
    for b, n_b in enumerate(sizes):
        m = 3 * n_b
        es.append(offs[b] + rs.randint(0, n_b, m))
        et.append(offs[b] + rs.randint(0, n_b, m))
    esrc = torch.from_numpy(np.concatenate(es).astype(np.int64))
    etgt = torch.from_numpy(np.concatenate(et).astype(np.int64))
    nE = int(esrc.numel())
    Etgt = torch.zeros(nN, nE)
    Etgt[etgt, torch.arange(nE)] = 1.0
    nf = rnd(403, nN, 13)
    ef = rnd(404, nE, 5)
    gy = rnd(405, len(sizes), 12)
    y = model(nf, ef, esrc, Etgt, batch)
    y.backward(gy)
    S.update({"m/nf": nf.numpy(), "m/ef": ef.numpy(), "m/esrc": esrc.numpy().astype(np.int32),
              "m/etgt": etgt.numpy().astype(np.int32), "m/gy": gy.numpy(), "m/out": y.detach().numpy(),
              **sd_np(model, "m/p/"), **{"m/grad/" + k: v.grad.numpy() for k, v in model.named_parameters() if v.grad is not None}})
    np.savez_compressed(os.path.join(HERE, "set2set_golden.npz"), **S)


# ---------------------------------------------------------------------------------------------------
# every model class of GCN/models.py, one GAT model family, Citeseer and Pubmed (VERDICT r01 missing #1, #2)
# ---------------------------------------------------------------------------------------------------

# (fixture key, class, constructor kwargs, solver method or None); nhid = 128 keeps GroupNorm well conditioned (4 channels
# per group); the reference's default nhid = 16 is the degenerate case of SURVEY F8 and is covered by gcn_golden.npz
GCN_MODEL_CASES = [
    ("GCN", "GCN", {}, None), ("RGCN2", "RGCN2", {}, None), ("GCN3norm", "GCN3norm", {}, None),
    ("RGCN3fullnorm", "RGCN3fullnorm", {}, None), ("ODEGCN3fullnorm_rk4", "ODEGCN3fullnorm", {}, "rk4"),
    ("GCNK4", "GCNK", {"nlayers": 4}, None), ("GCNKnorm4", "GCNKnorm", {"nlayers": 4}, None),
    ("RESK1_4", "RESK1", {"nlayers": 4}, None), ("RESK2_6", "RESK2", {"nlayers": 6}, None),
    ("RESK_5_r2", "RESK", {"nlayers": 5, "residue_layers": 2}, None),
    ("RESK_6_r3", "RESK", {"nlayers": 6, "residue_layers": 3}, None),
    ("RESK1norm4", "RESK1norm", {"nlayers": 4}, None), ("RESK2norm6", "RESK2norm", {"nlayers": 6}, None),
    ("RESKnorm_5_r2", "RESKnorm", {"nlayers": 5, "residue_layers": 2}, None),
    ("ODEK1_4_rk4", "ODEK1", {"nlayers": 4}, "rk4"), ("ODEK1_3_dopri5", "ODEK1", {"nlayers": 3}, "dopri5"),
    ("ODEK2_5_rk4", "ODEK2", {"nlayers": 5}, "rk4"),
    # the reference passes `dropout` as the ODE block's tolerance (GCN/models.py:587): dopri5 at rtol = atol = 0.5
    ("ODEK2_4_dopri5", "ODEK2", {"nlayers": 4}, "dopri5"),
]
DATASET_CASES = [  # (dataset, fixture key, class, kwargs, method, nhid)
    ("citeseer", "ODEGCN3_rk4", "ODEGCN3", {}, "rk4", 128), ("citeseer", "ODEGCN3_dopri5", "ODEGCN3", {}, "dopri5", 128),
    ("citeseer", "RGCN3", "RGCN3", {}, None, 128), ("citeseer", "ODEGCN3_rk4_h16", "ODEGCN3", {}, "rk4", 16),
    ("pubmed", "ODEGCN3_rk4", "ODEGCN3", {}, "rk4", 128), ("pubmed", "RGCN3", "RGCN3", {}, None, 128),
    ("pubmed", "RESK1_4", "RESK1", {"nlayers": 4}, None, 128), ("pubmed", "ODEK1_3_dopri5", "ODEK1", {"nlayers": 3}, "dopri5", 64),
]
GAT_MODEL_CASES = [("RGCN3", "RGCN3", {}, None), ("ODEGCN3_rk4", "ODEGCN3", {}, "rk4"), ("RESK1_4", "RESK1", {"nlayers": 4}, None)]


def _run_model_case(models_mod, cls_name, kw, method, nfeat, nhid, nclass, inputs, labels, idx_train, G_, key):
    """One model, twice: in float32 (THE fixture) and in float64.  The distance between the reference's own two results
    measures how ill-conditioned the case is (ReLU masks and GroupNorm groups near their singular points amplify fp32
    rounding); the tests hold the CUDA path to 1e-5 PLUS four times that distance, so a well-conditioned case is held to
    the bare fp32 bar and no case is asked for more digits than the reference itself has."""
    from tests import _golden as TG

    def run(dtype):
        stats = {}
        if method:
            set_method(models_mod, method, None, stats)
        model = getattr(models_mod, cls_name)(nfeat=nfeat, nhid=nhid, nclass=nclass, dropout=0.5, **kw)
        TG.fill_params(model)
        model = model.to(dtype).eval()
        for m in model.modules():                      # ODEBlock keeps its time grid as a float32 buffer-less tensor
            if hasattr(m, "integration_time"):
                m.integration_time = m.integration_time.to(dtype)
        ins = [t.to(dtype) if t.is_floating_point() else t for t in inputs]
        has_ode = method is not None
        if has_ode:
            model.nfe = 0
        out = model(*ins)
        nfe_f = int(model.nfe) if has_ode else 0
        if has_ode:
            model.nfe = 0
        loss = torch.nn.functional.nll_loss(out[idx_train], labels[idx_train])
        loss.backward()
        nfe_b = int(model.nfe) if has_ode else 0
        return model, out.detach(), loss.item(), nfe_f, nfe_b, stats

    model, out, loss, nfe_f, nfe_b, stats = run(torch.float32)
    model64, out64, loss64, _, _, stats64 = run(torch.float64)
    G_.update({key + "out": out.numpy(), key + "loss": np.float32(loss),
               key + "cond/out": np.float64((out.double() - out64).abs().max().item()),
               key + "cond/loss": np.float64(abs(loss - loss64)),
               key + "nfe_f": np.int64(nfe_f), key + "nfe_b": np.int64(nfe_b),
               key + "acc_f": np.int64(stats.get("forward", {}).get("accepted", 0)),
               key + "rej_f": np.int64(stats.get("forward", {}).get("rejected", 0)),
               key + "acc_b": np.int64(stats.get("backward", {}).get("accepted", 0)),
               key + "rej_b": np.int64(stats.get("backward", {}).get("rejected", 0)),
               # the same counts of the reference's float64 run: where they differ from the float32 ones the step sequence of
               # the case is not determined by the arithmetic, and the tests accept anything between the two (+-1)
               key + "acc_f64": np.int64(stats64.get("forward", {}).get("accepted", 0)),
               key + "rej_f64": np.int64(stats64.get("forward", {}).get("rejected", 0)),
               key + "acc_b64": np.int64(stats64.get("backward", {}).get("accepted", 0)),
               key + "rej_b64": np.int64(stats64.get("backward", {}).get("rejected", 0))})
    worst = 0.0
    p64 = dict(model64.named_parameters())
    for pn, p_ in model.named_parameters():
        if p_.grad is None:
            continue
        G_[key + "grad/" + pn] = TG.grad_sample(p_.grad).numpy().copy()
        G_[key + "gradnorm/" + pn] = np.float64(p_.grad.double().norm().item())
        G_[key + "gradmax/" + pn] = np.float64(p_.grad.abs().max().item())
        cond = float((p_.grad.double() - p64[pn].grad).abs().max().item())
        G_[key + "cond/" + pn] = np.float64(cond)
        worst = max(worst, cond / max(float(p_.grad.abs().max()), 1e-30))
    print(key, "loss %.6f nfe %d/%d" % (loss, nfe_f, nfe_b), stats, stats64 if stats64 != stats else "", "| reference fp32 vs fp64: logits %.1e, worst gradient %.1e of its max" % (
        float(G_[key + "cond/out"]) / float(out.abs().max()), worst))


def make_models():
    """models_golden.npz: logits, loss, NFE / step counts and (sampled) parameter gradients of every model class of the
    reference on Cora, plus Citeseer (real features) and Pubmed (real graph, synthetic features) cases and a GAT family.
    Parameters are not stored: both sides fill them by name with tests/_golden.py:fill_params."""
    _install_shims()
    from tests import _golden as TG
    gcn = load_ref("GCN")
    gat = load_ref("GAT")
    M = {}
    data = {}
    for ds in ("cora", "citeseer", "pubmed"):
        c = np.load(os.path.join(HERE, "planetoid_%s.npz" % ds))
        n = int(c["n"])
        adj = TG.coo_adj(c["coo_row"], c["coo_col"], c["coo_val"], n)
        if ds != "pubmed":
            import scipy.sparse as sp
            m = sp.csr_matrix((c["feat_data"], c["feat_indices"], c["feat_indptr"]), shape=(n, int(c["nfeat"])))
            feats = torch.from_numpy(np.asarray(m.todense(), dtype=np.float32))
            labels = torch.from_numpy(c["labels"].astype(np.int64))
            idx = torch.from_numpy(c["idx_train"].astype(np.int64))
        else:
            feats = TG.pubmed_features(n)
            labels = torch.from_numpy(np.random.RandomState(1).randint(0, 3, n).astype(np.int64))
            idx = torch.arange(60)
        data[ds] = (c, n, adj, feats, labels, idx)

    c, n, adj, feats, labels, idx = data["cora"]
    for key, cls_name, kw, method in GCN_MODEL_CASES:
        _run_model_case(gcn.models, cls_name, kw, method, feats.shape[1], 128, 7, (feats, adj), labels, idx, M, "cora/%s/" % key)
    for ds, key, cls_name, kw, method, nhid in DATASET_CASES:
        c, n, adj, feats, labels, idx = data[ds]
        _run_model_case(gcn.models, cls_name, kw, method, feats.shape[1], nhid, int(labels.max()) + 1, (feats, adj), labels, idx, M,
                        "%s/%s/" % (ds, key))
    # GAT family on the Cora edge list (GAT/utils.py:187-196: each undirected edge once)
    c, n, adj, feats, labels, idx = data["cora"]
    src = torch.from_numpy(c["gat_src"].astype(np.int64))
    tgt = torch.from_numpy(c["gat_tgt"].astype(np.int64))
    E = len(src)
    Mtgt = torch.sparse_coo_tensor(torch.stack([tgt, torch.arange(E)]), torch.ones(E), (n, E))
    for key, cls_name, kw, method in GAT_MODEL_CASES:
        _run_model_case(gat.models, cls_name, kw, method, feats.shape[1], 128, 7, (feats, src, tgt, Mtgt), labels, idx, M,
                        "gat_cora/%s/" % key)
    np.savez_compressed(os.path.join(HERE, "models_golden.npz"), **M)
    print("models_golden.npz", os.path.getsize(os.path.join(HERE, "models_golden.npz")) // 1024, "KiB")


QC_MODEL_CASES = [  # (fixture key, class in QC/layer_models.py, hidden, num_layers)
    ("MPNN_ENN_K_Sum", "MPNN_ENN_K_Sum", 24, 3), ("MPNN_ENN_K_Set2Set", "MPNN_ENN_K_Set2Set", 24, 3),
    ("EdgeRES1_K_Set2Set", "EdgeRES1_K_Set2Set", 64, 4), ("EdgeGCN_K_Sum_h73", "EdgeGCN_K_Sum", 73, 3),
]


def qc_batch(sizes=(7, 1, 12, 3, 18, 9, 5, 22), seed=21):
    """A block-diagonal batch of molecule-shaped graphs (node ids offset per molecule), numpy-seeded so that the tests
    rebuild it: node features [N, 13], edge features [E, 5], esrc / etgt [E], batch [N]."""
    rs = np.random.RandomState(seed)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    es, et = [], []
    for b, n_b in enumerate(sizes):
        m = max(2 * n_b, 1)
        es.append(offs[b] + rs.randint(0, n_b, m))
        et.append(offs[b] + rs.randint(0, n_b, m))
    esrc = torch.from_numpy(np.concatenate(es).astype(np.int64))
    etgt = torch.from_numpy(np.concatenate(et).astype(np.int64))
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    nN, nE = int(offs[-1]), int(esrc.numel())
    return rnd(seed + 1, nN, 13), rnd(seed + 2, nE, 5), esrc, etgt, batch, rnd(seed + 3, len(sizes), 12)


def make_qc_models():
    """qc_models_golden.npz: the QC model classes that had no reference fixture:
    ``MPNN_enn_edge`` (QC/mpnn.py:5-32) and, from QC/layer_models.py, ``MPNN_ENN_K_Sum``, ``MPNN_ENN_K_Set2Set``,
    ``EdgeRES1_K_Set2Set`` (hidden 64: GroupNorm(32, 73) cannot be built) and ``EdgeGCN_K_Sum`` at the reference's default
    hidden = 73.  Parameters by name (fill_params), eval mode, float32 fixture + float64 conditioning run."""
    _install_shims()
    from tests import _golden as TG
    qc = load_ref("QC", names=("mpnn", "layer_models"))
    Q = {}
    nf, ef, esrc, etgt, batch, gy = qc_batch()
    nN, nE = nf.shape[0], ef.shape[0]
    Etgt = torch.zeros(nN, nE)
    Etgt[etgt, torch.arange(nE)] = 1.0

    def grads(model, model64, key):
        p64 = dict(model64.named_parameters())
        for pn, p_ in model.named_parameters():
            if p_.grad is None:
                continue
            Q[key + "grad/" + pn] = TG.grad_sample(p_.grad).numpy().copy()
            Q[key + "gradnorm/" + pn] = np.float64(p_.grad.double().norm().item())
            Q[key + "gradmax/" + pn] = np.float64(p_.grad.abs().max().item())
            Q[key + "cond/" + pn] = np.float64((p_.grad.double() - p64[pn].grad).abs().max().item())

    # the message-passing core alone: x [N, h], edge matrices [E, h, h]
    h = 24
    outs = {}
    for dt in (torch.float32, torch.float64):
        net = qc.mpnn.MPNN_enn_edge(5, h)
        net.set_T(3)
        TG.fill_params(net)
        net = net.to(dt)
        x = rnd(31, nN, h).to(dt).requires_grad_(True)
        ed = rnd(32, nE, h, h, scale=1.0 / h ** 0.5).to(dt).requires_grad_(True)
        y = net(x, esrc, Etgt.to_sparse().to(dt), ed)
        y.backward(rnd(33, nN, h).to(dt))
        outs[dt] = (net, y.detach(), x.grad, ed.grad)
    net, y, gx, ged = outs[torch.float32]
    net64, y64, gx64, ged64 = outs[torch.float64]
    Q.update({"mpnn/out": y.numpy(), "mpnn/grad_x": gx.numpy(), "mpnn/grad_ed": TG.grad_sample(ged).numpy().copy(),
              "mpnn/cond/out": np.float64((y.double() - y64).abs().max().item()),
              "mpnn/cond/grad_x": np.float64((gx.double() - gx64).abs().max().item()),
              "mpnn/cond/grad_ed": np.float64((ged.double() - ged64).abs().max().item())})
    grads(net, net64, "mpnn/")
    print("mpnn/", "reference fp32 vs fp64: out %.1e" % float(Q["mpnn/cond/out"]))

    for key, cls_name, hidden, K in QC_MODEL_CASES:
        res = {}
        for dt in (torch.float32, torch.float64):
            model = getattr(qc.layer_models, cls_name)(node_features=13, edge_features=5, target_features=12, hidden_features=hidden,
                                                       num_layers=K, s2s_processing_steps=3, type="regression", dropout=0.0)
            TG.fill_params(model)
            model = model.to(dt).eval()
            out = model(nf.to(dt), ef.to(dt), esrc, Etgt.to(dt), batch)
            out.backward(gy.to(dt))
            res[dt] = (model, out.detach())
        model, out = res[torch.float32]
        model64, out64 = res[torch.float64]
        k = "m/%s/" % key
        Q.update({k + "out": out.numpy(), k + "cond/out": np.float64((out.double() - out64).abs().max().item())})
        grads(model, model64, k)
        print(k, "out max %.3f, reference fp32 vs fp64 %.1e" % (float(out.abs().max()), float(Q[k + "cond/out"])))
    np.savez_compressed(os.path.join(HERE, "qc_models_golden.npz"), **Q)
    print("qc_models_golden.npz", os.path.getsize(os.path.join(HERE, "qc_models_golden.npz")) // 1024, "KiB")


def make_qc_collate():
    """qc_collate_golden.npz: the reference's DataLoader collate (QC/datasets/utils.py:153-217) on synthetic molecules.
    ``QC/datasets/utils.py`` imports rdkit and networkx at module level (utils.py:15-19) for its file parsers; empty stub
    modules satisfy the import, the collate itself is numpy + torch."""
    from tests import _golden as TG
    for name in ("rdkit", "networkx"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    spec = importlib.util.spec_from_file_location("ref_qc_datasets_utils", os.path.join(REF, "QC", "datasets", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {}
    for tag, kw in (("a", {}), ("b", {"sizes": (3, 5, 1, 1, 8), "seed": 5})):
        mols = TG.synthetic_molecules(**kw)
        bs, G_, B, X, E_d, E_src, E_tgt, Y = mod.collate_g_concat_edge_data(mols)
        out.update({tag + "/bs": np.int64(bs), tag + "/G": G_.numpy(), tag + "/B": B.numpy(), tag + "/X": X.numpy(),
                    tag + "/E_d": E_d.numpy(), tag + "/E_src": E_src.numpy(), tag + "/E_tgt": E_tgt.numpy(), tag + "/Y": Y.numpy()})
    np.savez_compressed(os.path.join(HERE, "qc_collate_golden.npz"), **out)
    print("qc_collate_golden.npz", os.path.getsize(os.path.join(HERE, "qc_collate_golden.npz")) // 1024, "KiB")


def orbit_problem(n_sys=3, bodies=4):
    """Relation structure of ``prototypes/orbit/train_IN.py:process_instance`` (every ordered pair of distinct bodies of a
    system is a relation) for a block-diagonal batch: (n, m, src[m], tgt[m])."""
    src, tgt = [], []
    for s_ in range(n_sys):
        for a in range(bodies):
            for b in range(bodies):
                if a != b:
                    src.append(s_ * bodies + a)
                    tgt.append(s_ * bodies + b)
    return n_sys * bodies, len(src), np.array(src, dtype=np.int64), np.array(tgt, dtype=np.int64)


def make_orbit():
    """orbit_golden.npz: the reference's Interaction Network and IN-ODE (prototypes/orbit/model.py:29-171) on a small batch
    of n-body systems -- outputs, input gradients and parameter gradients, the ODE model with the fixed-step rk4 solver
    (arithmetic pinned step by step) and with the reference's default dopri5 (NFE and step counts as well)."""
    _install_shims()
    spec = importlib.util.spec_from_file_location("ref_orbit_model", os.path.join(REF, "prototypes", "orbit", "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    n, m, src, tgt = orbit_problem()
    Msrc = torch.zeros(n, m)
    Mtgt = torch.zeros(n, m)
    Msrc[torch.from_numpy(src), torch.arange(m)] = 1
    Mtgt[torch.from_numpy(tgt), torch.arange(m)] = 1
    out = {"n": np.int64(n), "m": np.int64(m), "src": src, "tgt": tgt}
    O = rnd(301, n, 5)
    G = rnd(302, n, 2)

    def grads(model, key, fwd):
        model.zero_grad()
        o = O.clone().requires_grad_(True)
        P = fwd(model, o)
        (P * G).sum().backward()
        out[key + "/P"] = P.detach().numpy()
        out[key + "/grad_O"] = o.grad.numpy()
        for k, p_ in model.named_parameters():
            if p_.grad is not None:
                out[key + "/grad/" + k] = p_.grad.numpy().copy()

    torch.manual_seed(31)
    net = mod.IN(5, 0, 0, 2)
    out.update(sd_np(net, "IN/sd/"))
    grads(net, "IN", lambda mdl, o: mdl(o, None, None, Msrc, Mtgt))

    torch.manual_seed(32)
    ode = mod.IN_ODE(5, 0, 0, 2)
    out.update(sd_np(ode, "IN_ODE/sd/"))
    for method in ("rk4", "dopri5"):
        stats = {}
        mod.odeint = functools.partial(restated.odeint_adjoint, method=method, stats=stats)
        ode.nfe = 0
        grads(ode, "IN_ODE_" + method, lambda mdl, o: mdl(o, None, None, Msrc, Mtgt))
        out["IN_ODE_%s/nfe" % method] = np.int64(ode.nfe)
        out["IN_ODE_%s/stats" % method] = np.array(json.dumps(stats))
    np.savez_compressed(os.path.join(HERE, "orbit_golden.npz"), **out)
    print("orbit_golden.npz", os.path.getsize(os.path.join(HERE, "orbit_golden.npz")) // 1024, "KiB",
          {k: v for k, v in out.items() if k.endswith("/nfe") or k.endswith("/stats")})


if __name__ == "__main__":
    if "--only-orbit" in sys.argv:
        make_orbit()
    elif "--only-qc-collate" in sys.argv:
        make_qc_collate()
    elif "--only-qc-models" in sys.argv:
        make_qc_models()
    elif "--only-set2set" in sys.argv:
        make_set2set()
    elif "--only-models" in sys.argv:
        make_models()
    else:
        main()
        make_models()
        make_qc_models()
        make_qc_collate()
        make_orbit()
