"""GPU parity of EVERY model class of the reference (GCN/models.py:8-600, one GAT family from GAT/models.py) and of the
Citeseer / Pubmed datasets against fixtures produced by the unmodified reference modules (tests/golden/make_golden.py
``make_models`` -> models_golden.npz).  VERDICT r01 missing #1 / #2.

Parameters are not stored in the fixture: both sides fill them by NAME with ``tests/_golden.py:fill_params`` -- so a
parameter that exists under a different name (or shape) on one side fails immediately, which pins the ``state_dict``
surface as well.  Models run in eval mode (dropout off; the CPU and CUDA dropout streams differ -- SURVEY 8c(5)).

Bars: logits, loss and parameter gradients (a strided sample of at most 4096 elements plus the full tensor's norm) within
1e-5 of the tensor's scale PLUS four times the distance between the reference's OWN float32 and float64 results for that
tensor (stored in the fixture: well-conditioned cases are held to the bare fp32 bar, and no case is asked for more digits
than the reference itself has).  Adaptive (dopri5) cases: accepted step counts exact, values to the solver's tolerance.
"""
import numpy as np
import pytest
import torch

from tests import _golden as G
from tests.golden.make_golden import DATASET_CASES, GAT_MODEL_CASES, GCN_MODEL_CASES

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = dict(rtol=1e-5, atol_scale=1e-5)


def _dataset(ds):
    c = G.load("planetoid_" + ds)
    n = int(c["n"])
    adj = G.coo_adj(c["coo_row"], c["coo_col"], c["coo_val"], n)
    if ds != "pubmed":
        feats = G.dense_features(ds)
        labels = torch.from_numpy(c["labels"].astype(np.int64))
        idx = torch.from_numpy(c["idx_train"].astype(np.int64))
    else:
        feats = G.pubmed_features(n)
        labels = torch.from_numpy(np.random.RandomState(1).randint(0, 3, n).astype(np.int64))
        idx = torch.arange(60)
    return c, n, adj, feats, labels, idx


def _run(models_mod, cls_name, kw, method, nhid, inputs, labels, idx, key):
    g = G.load("models_golden")
    nfeat, nclass = inputs[0].shape[1], int(labels.max()) + 1
    model = getattr(models_mod, cls_name)(nfeat=nfeat, nhid=nhid, nclass=nclass, dropout=0.5, **kw)
    G.fill_params(model)
    model = model.to(DEV).eval()
    blocks = [m for m in model.modules() if isinstance(m, models_mod.ODEBlock)]
    for b in blocks:
        b.method = method
        b.stats = {}
        b.nfe = 0
    out = model(*[t.to(DEV) for t in inputs])
    nfe_f = sum(b.nfe for b in blocks)
    for b in blocks:
        b.nfe = 0
    loss = torch.nn.functional.nll_loss(out[idx.to(DEV)], labels.to(DEV)[idx.to(DEV)])
    loss.backward()
    nfe_b = sum(b.nfe for b in blocks)
    adaptive = method == "dopri5"
    if blocks:
        for phase, a_key, r_key in (("forward", "acc_f", "rej_f"), ("backward", "acc_b", "rej_b")):
            acc = sum(b.stats.get(phase, {}).get("accepted", 0) for b in blocks)
            rej = sum(b.stats.get(phase, {}).get("rejected", 0) for b in blocks)
            w32 = (int(g[key + a_key]), int(g[key + r_key]))
            w64 = (int(g[key + a_key + "64"]), int(g[key + r_key + "64"]))
            if w32 == w64:      # the reference's own float32 and float64 runs agree on the step sequence: the ACCEPTED steps must
                # be reproduced exactly (north_star: "adaptive solvers compared on the accepted-step count"); a trial step whose
                # error ratio sits within rounding of 1.0 may be rejected on one side only, hence +-1 on the rejected count
                assert acc == w32[0] and abs(rej - w32[1]) <= 1, (phase, (acc, rej), w32)
            else:               # the case's step sequence depends on rounding: anything between the two runs (+-1) is the reference's
                for got_, a_, b_ in ((acc, w32[0], w64[0]), (rej, w32[1], w64[1])):
                    assert min(a_, b_) - 1 <= got_ <= max(a_, b_) + 1, (phase, (acc, rej), w32, w64)
        if int(g[key + "nfe_f"]) and not adaptive:     # (the reference's K-layer ODE models have no nfe property: 0 in the fixture)
            assert (nfe_f, nfe_b) == (int(g[key + "nfe_f"]), int(g[key + "nfe_b"]))
    # Bar: 1e-5 (adaptive: the solver's own 1e-4) of the tensor's scale PLUS four times the distance between the reference's
    # own float32 and float64 results for that tensor (fixture key cond/...): well-conditioned cases are held to the bare
    # fp32 bar, and no case is asked for more digits than the reference itself has.
    bar = 1e-5 if not adaptive else 1e-4
    want = torch.from_numpy(g[key + "out"]).double()
    err = float((out.detach().cpu().double() - want).abs().max())
    allowed = bar * float(want.abs().max()) + 4 * float(g[key + "cond/out"])
    assert err <= allowed, "%slogits: max err %.3e > %.3e (reference fp32-vs-fp64 %.1e)" % (key, err, allowed, float(g[key + "cond/out"]))
    assert abs(float(loss.detach()) - float(g[key + "loss"])) <= bar * max(1.0, abs(float(g[key + "loss"]))) + 4 * float(g[key + "cond/loss"])
    names = [k[len(key + "grad/"):] for k in g if k.startswith(key + "grad/")]
    assert sorted(names) == sorted(n_ for n_, p in model.named_parameters() if p.grad is not None), "parameter names differ"
    worst, worst_cond = 0.0, 0.0
    for pn, p in model.named_parameters():
        if p.grad is None:
            continue
        want = torch.from_numpy(g[key + "grad/" + pn]).double()
        got = G.grad_sample(p.grad).cpu().double()
        scale = float(g[key + "gradmax/" + pn])
        cond = float(g[key + "cond/" + pn])
        diff = (got - want).abs()
        err = float(diff.max())
        allowed = bar * scale + 4 * cond + 1e-9
        n_bad = int((diff > allowed).sum())
        # BULK at the bar: at least 95 % of a tensor's sampled entries; every entry within 100 x the bar (+ 8 x the reference's
        # own fp32-vs-fp64 distance).  One ReLU mask element that falls on the other side of zero moves a handful of entries by
        # a discrete amount (tests/test_gpu_gcn.py::test_relu_regime_gradient_parity measures and bounds that effect); the 1e-9
        # floor is for gradients that are identically zero in exact arithmetic (the attention bias under the softmax).
        assert n_bad <= max(0.05 * diff.numel(), 4) and err <= 100 * bar * scale + 8 * cond + 1e-9, \
            "%s%s: max err %.3e > %.3e (scale %.3e, reference fp32-vs-fp64 %.1e), %d/%d outside" % (key, pn, err, allowed, scale, cond, n_bad, diff.numel())
        norm = float(p.grad.double().norm())
        wn = float(g[key + "gradnorm/" + pn])
        assert abs(norm - wn) <= 100 * bar * wn + 8 * cond * p.grad.numel() ** 0.5 + 1e-9, (key, pn, "norm", norm, wn)
        if scale > 1e-7:
            worst, worst_cond = max(worst, err / scale), max(worst_cond, cond / scale)
    print("%s ok: worst gradient error %.1e of the tensor's max (reference fp32-vs-fp64: %.1e)" % (key, worst, worst_cond))


@pytest.mark.parametrize("case", GCN_MODEL_CASES, ids=[c[0] for c in GCN_MODEL_CASES])
def test_gcn_models_cora(case):
    import graph_odenet_b200  # noqa: F401
    from graph_odenet_b200.GCN import models
    key, cls_name, kw, method = case
    c, n, adj, feats, labels, idx = _dataset("cora")
    _run(models, cls_name, kw, method, 128, (feats, adj), labels, idx, "cora/%s/" % key)


@pytest.mark.parametrize("case", DATASET_CASES, ids=["%s-%s" % (c[0], c[1]) for c in DATASET_CASES])
def test_gcn_models_citeseer_pubmed(case):
    import graph_odenet_b200  # noqa: F401
    from graph_odenet_b200.GCN import models
    ds, key, cls_name, kw, method, nhid = case
    c, n, adj, feats, labels, idx = _dataset(ds)
    _run(models, cls_name, kw, method, nhid, (feats, adj), labels, idx, "%s/%s/" % (ds, key))


@pytest.mark.parametrize("case", GAT_MODEL_CASES, ids=[c[0] for c in GAT_MODEL_CASES])
def test_gat_models_cora(case):
    import graph_odenet_b200  # noqa: F401
    from graph_odenet_b200.GAT import models
    key, cls_name, kw, method = case
    c, n, adj, feats, labels, idx = _dataset("cora")
    src = torch.from_numpy(c["gat_src"].astype(np.int64))
    tgt = torch.from_numpy(c["gat_tgt"].astype(np.int64))
    E = len(src)
    Mtgt = torch.sparse_coo_tensor(torch.stack([tgt, torch.arange(E)]), torch.ones(E), (n, E))
    _run(models, cls_name, kw, method, 128, (feats, src, tgt, Mtgt), labels, idx, "gat_cora/%s/" % key)
