"""GPU parity of EVERY model class of the reference (GCN/models.py:8-600, one GAT family from GAT/models.py) and of the
Citeseer / Pubmed datasets against fixtures produced by the unmodified reference modules (tests/golden/make_golden.py
``make_models`` -> models_golden.npz).  VERDICT r01 missing #1 / #2.

Parameters are not stored in the fixture: both sides fill them by NAME with ``tests/_golden.py:fill_params`` -- so a
parameter that exists under a different name (or shape) on one side fails immediately, which pins the ``state_dict``
surface as well.  Models run in eval mode (dropout off; the CPU and CUDA dropout streams differ -- SURVEY 8c(5)).

Bars: logits and loss 1e-5; parameter gradients 1e-5 of the tensor's scale, stored as a strided sample of at most 4096
elements plus the full tensor's norm.  Adaptive (dopri5) cases: accepted / rejected step counts exact, values to the
solver's own tolerance.  The exceptions are stated where they are asserted.
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
            assert (acc, rej) == (int(g[key + a_key]), int(g[key + r_key])), (phase, acc, rej, int(g[key + a_key]), int(g[key + r_key]))
        if int(g[key + "nfe_f"]):     # (the reference's K-layer ODE models have no nfe property: the fixture holds 0 there)
            assert (nfe_f, nfe_b) == (int(g[key + "nfe_f"]), int(g[key + "nfe_b"]))
    # adaptive solves agree to the solver's tolerance, not to fp32 rounding (see tests/test_gpu_gcn.py::test_ode_block_golden)
    vtol = TOL if not adaptive else dict(rtol=1e-4, atol_scale=1e-4)
    G.assert_close(out, g[key + "out"], **vtol, what=key + "logits")
    assert abs(float(loss.detach()) - float(g[key + "loss"])) < (1e-5 if not adaptive else 1e-4) * max(1.0, abs(float(g[key + "loss"])))
    names = [k[len(key + "grad/"):] for k in g if k.startswith(key + "grad/")]
    assert sorted(names) == sorted(n_ for n_, p in model.named_parameters() if p.grad is not None), "parameter names differ"
    worst = 0.0
    for pn, p in model.named_parameters():
        if p.grad is None:
            continue
        want = torch.from_numpy(g[key + "grad/" + pn])
        got = G.grad_sample(p.grad).cpu()
        scale = float(g[key + "gradnorm/" + pn]) / max(p.grad.numel(), 1) ** 0.5      # rms of the full tensor
        err = float((got.double() - want.double()).abs().max())
        worst = max(worst, err / max(scale, 1e-30))
        # gradient bar: 1e-5 relative + 1e-5 of the tensor's rms magnitude (adaptive: 1e-3)
        gt = 1e-5 if not adaptive else 1e-3
        bad = (got.double() - want.double()).abs() > gt * want.double().abs() + gt * max(scale, float(want.abs().max()))
        assert not bool(bad.any()), "%s%s: max err %.3e vs rms %.3e (%d/%d beyond %.0e)" % (key, pn, err, scale, int(bad.sum()), bad.numel(), gt)
        norm = float(p.grad.double().norm())
        assert abs(norm - float(g[key + "gradnorm/" + pn])) <= 10 * gt * float(g[key + "gradnorm/" + pn]) + 1e-12, (key, pn, "norm")
    print("%s ok: worst gradient error / rms %.2e" % (key, worst))


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
