"""GPU parity of the GAT path (gode_gat_fwd / gode_gat_bwd behind graph-odenet_b200/GAT) against the golden
fixtures produced by the unmodified reference (GAT/layers.py, GAT/models.py) and against the CPU oracle.

Bar: fp32, 1e-5 relative (+1e-5 of the tensor's max magnitude as the absolute floor)."""
import numpy as np
import pytest
import torch

from oracle import gat_ref, odeint as oracle_odeint
from tests import _golden as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pkg():
    import graph_odenet_b200  # noqa: F401
    from graph_odenet_b200 import ops
    from graph_odenet_b200.GAT import layers, models
    return ops, layers, models


def _cora_edges():
    c = G.load("planetoid_cora")
    return (torch.from_numpy(c["gat_src"].astype(np.int64)), torch.from_numpy(c["gat_tgt"].astype(np.int64)), int(c["n"]))


def test_gat_layer_golden():
    ops, layers, _ = _pkg()
    g = G.load("gat_golden")
    src, tgt, n = _cora_edges()
    lay = layers.GraphConvolution(16, 8).to(DEV)
    lay.load_state_dict({k: v for k, v in G.params(g, "gc/p/").items()})
    x = G.rnd(61, n, 16).to(DEV).requires_grad_(True)
    y = lay(x, src.to(DEV), tgt.to(DEV), None)
    y.backward(G.rnd(62, n, 8).to(DEV))
    G.assert_close(y, g["gc/out"], rtol=1e-5, atol_scale=1e-5, what="out")
    G.assert_close(x.grad, g["gc/grad_x"], rtol=1e-5, atol_scale=1e-5, what="grad_x")
    for k, p in lay.named_parameters():
        # w.bias shifts every a_e equally and a - max(a) is shift-invariant: its gradient is mathematically 0 (the
        # reference's autograd gets an exact 0 by summing and negating the same numbers); here two differently
        # ordered fp32 sums cancel to ~1e-6 of the summed magnitudes -> absolute floor
        G.assert_close(p.grad, g["gc/grad/" + k], rtol=1e-5, atol_scale=1e-5, what=k, atol_abs=1e-5 if k == "w.bias" else 0.0)
    # nodes without incoming edges output exactly 0 (679 of them on Cora)
    indeg = torch.bincount(tgt, minlength=n)
    assert int((indeg == 0).sum()) == 679
    assert float(y[(indeg == 0).to(DEV)].abs().max()) == 0.0


def test_gat_odefunc_golden():
    _, _, models = _pkg()
    g = G.load("gat_golden")
    src, tgt, n = _cora_edges()
    f = models.ODEfunc(16).to(DEV)
    f.load_state_dict(G.params(g, "odefunc/p/"))
    f.set_adj(src.to(DEV), tgt.to(DEV), None)
    x = G.rnd(63, n, 16).to(DEV).requires_grad_(True)
    t = torch.tensor(0.25, device=DEV, requires_grad=True)
    y = f(t, x)
    grads = torch.autograd.grad(y, (x, t) + tuple(f.parameters()), G.rnd(64, n, 16).to(DEV))
    # hidden=16: GroupNorm(16 groups of ONE channel) outputs beta + rounding noise, so the layer's input is
    # noise-level sensitive to how x*(gamma*rstd) + (beta - mean*gamma*rstd) is rounded: absolute tolerances here, as in
    # tests/test_gpu_gcn.py::test_odefunc_golden[16]; the non-degenerate widths are held to 1e-5 in test_gat_matches_oracle
    assert float((y.detach().cpu() - torch.from_numpy(g["odefunc/out"])).abs().max()) < 1e-4
    assert float((grads[0].cpu() - torch.from_numpy(g["odefunc/grad_x"])).abs().max()) < 2e-2
    G.assert_close(grads[1], g["odefunc/grad_t"], rtol=1e-3, atol_abs=1e-5, what="grad_t")
    for (k, _), gr in zip(f.named_parameters(), grads[2:]):
        if "norm1.weight" in k:
            continue                                           # sum of rounding noise (x - mean == 0 exactly)
        tol = dict(rtol=1e-4, atol_scale=1e-4)
        if k.endswith("w.bias"):
            tol = dict(rtol=1e-5, atol_scale=1e-5, atol_abs=1e-5)      # mathematically zero, see test_gat_layer_golden
        if k.endswith("w.weight"):
            tol = dict(rtol=1e-5, atol_scale=1e-5, atol_abs=1e-6)      # input rows are all ~beta: the reference value is 6e-9 of noise
        G.assert_close(gr, g["odefunc/grad/" + k], what=k, **tol)
    assert f.nfe == 1


def _random_edges(n, e, seed, isolated=5):
    rs = np.random.RandomState(seed)
    src = rs.randint(0, n, e)
    tgt = rs.randint(isolated, n, e)            # nodes [0, isolated) have no incoming edge
    tgt[: e // 8] = n - 1                       # one hub with many incoming edges
    src[-3:] = src[-6:-3]
    tgt[-3:] = tgt[-6:-3]                       # duplicate edges are separate terms of the sums
    return torch.from_numpy(src.astype(np.int64)), torch.from_numpy(tgt.astype(np.int64))


@pytest.mark.parametrize("vec", ["1", "0"], ids=["lane-group", "scalar"])
@pytest.mark.parametrize("i,o,heads", [(16, 7, 1), (9, 16, 1), (33, 16, 8), (12, 3, 2), (20, 64, 1), (10, 32, 1), (14, 8, 4),
                                       (11, 128, 1), (13, 4, 4)])
def test_gat_matches_oracle(i, o, heads, vec, monkeypatch):
    """Single- and multi-head layers against the oracle (H independent reference heads, concatenated), on the lane-group
    kernels (128-bit rows; widths 16 / 32 / 64 / 128 with power-of-two heads) and on the scalar kernels (GODE_GAT_VEC=0,
    and every other width)."""
    monkeypatch.setenv("GODE_GAT_VEC", vec)
    ops, layers, _ = _pkg()
    n, e = 700, 5000
    src, tgt = _random_edges(n, e, seed=i + o)
    torch.manual_seed(i * 100 + o)
    lay = layers.GraphConvolution(i, o, heads=heads)
    with torch.no_grad():
        lay.f.bias.uniform_(-0.3, 0.3)
        lay.w.bias.uniform_(-0.3, 0.3)
    x = torch.randn(n, i)
    gy = torch.randn(n, o * heads)
    # oracle: per-head reference layers
    pc = {k: v.detach().clone().requires_grad_(True) for k, v in lay.state_dict().items()}
    xo = x.clone().requires_grad_(True)
    hs = [(pc["f.weight"][h * o:(h + 1) * o], pc["f.bias"][h * o:(h + 1) * o], pc["w.weight"][h:h + 1], pc["w.bias"][h:h + 1])
          for h in range(heads)]
    yo = gat_ref.gat_multihead(xo, src, tgt, hs)
    yo.backward(gy)
    lay = lay.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    y = lay(xg, src.to(DEV), tgt.to(DEV), None)
    y.backward(gy.to(DEV))
    G.assert_close(y, yo, rtol=1e-5, atol_scale=1e-5, what="out")
    G.assert_close(xg.grad, xo.grad, rtol=1e-5, atol_scale=1e-5, what="grad_x")
    for k, p in lay.named_parameters():
        G.assert_close(p.grad, pc[k].grad, rtol=1e-5, atol_scale=2e-5, what=k, atol_abs=1e-5 if k == "w.bias" else 0.0)


@pytest.mark.parametrize("vec", ["1", "0"], ids=["lane-group", "scalar"])
@pytest.mark.parametrize("heads,o", [(1, 16), (8, 16)])
def test_gat_hub_nodes_are_chunked(heads, o, vec, monkeypatch):
    """Power-law hubs: nodes with more than GODE_GAT_CHUNK (64) in- or out-edges are reduced by several threads whose
    partial sums are added in chunk order; one hub has exactly 64 edges, one 65, one 1000 -- against the oracle."""
    monkeypatch.setenv("GODE_GAT_VEC", vec)
    ops, layers, _ = _pkg()
    n, i = 1500, 12
    rng = np.random.default_rng(5)
    src = [rng.integers(0, n, 3000)]
    tgt = [rng.integers(0, n, 3000)]
    for hub, deg in ((7, 64), (8, 65), (9, 1000)):
        src += [rng.integers(0, n, deg), np.full(deg, hub)]          # hub as target, then hub as source
        tgt += [np.full(deg, hub), rng.integers(0, n, deg)]
    src = torch.from_numpy(np.concatenate(src).astype(np.int64))
    tgt = torch.from_numpy(np.concatenate(tgt).astype(np.int64))
    perm = torch.from_numpy(rng.permutation(src.numel()))
    src, tgt = src[perm], tgt[perm]
    torch.manual_seed(3)
    lay = layers.GraphConvolution(i, o, heads=heads)
    x = torch.randn(n, i)
    gy = torch.randn(n, o * heads)
    pc = {k: v.detach().clone().requires_grad_(True) for k, v in lay.state_dict().items()}
    xo = x.clone().requires_grad_(True)
    hs = [(pc["f.weight"][h * o:(h + 1) * o], pc["f.bias"][h * o:(h + 1) * o], pc["w.weight"][h:h + 1], pc["w.bias"][h:h + 1])
          for h in range(heads)]
    yo = gat_ref.gat_multihead(xo, src, tgt, hs)
    yo.backward(gy)
    lay = lay.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    graph = ops.gat_graph_for(src.to(DEV), tgt.to(DEV), n)
    indeg, outdeg = torch.bincount(tgt, minlength=n), torch.bincount(src, minlength=n)
    assert graph.c.t_heavy.n_heavy == int((indeg > 64).sum()) >= 2
    assert graph.c.s_heavy.n_heavy == int((outdeg > 64).sum()) >= 2
    assert graph.c.t_heavy.n_chunks == int(((indeg[indeg > 64] + 63) // 64).sum())
    y = lay(xg, src.to(DEV), tgt.to(DEV), None)
    y.backward(gy.to(DEV))
    G.assert_close(y, yo, rtol=1e-5, atol_scale=1e-5, what="out")
    G.assert_close(xg.grad, xo.grad, rtol=1e-5, atol_scale=1e-5, what="grad_x")
    for k, p in lay.named_parameters():
        G.assert_close(p.grad, pc[k].grad, rtol=1e-5, atol_scale=2e-5, what=k, atol_abs=1e-5 if k == "w.bias" else 0.0)


def test_gat_empty_and_degenerate():
    ops, layers, _ = _pkg()
    lay = layers.GraphConvolution(4, 8).to(DEV)
    x = torch.randn(10, 4, device=DEV, requires_grad=True)
    e0 = torch.empty(0, dtype=torch.int64, device=DEV)
    y = lay(x, e0, e0, None)                      # no edges: every node outputs 0 / (0 + eps) = 0
    assert y.shape == (10, 8) and float(y.abs().max()) == 0.0
    y.sum().backward()
    assert float(x.grad.abs().max()) == 0.0
    with pytest.raises(IndexError):
        lay(x, torch.tensor([0, 11], device=DEV), torch.tensor([1, 2], device=DEV), None)
    with pytest.raises(TypeError):
        lay(x, torch.tensor([0]), torch.tensor([1]), None)      # CPU edge list: no CPU path
    xn = x.detach().clone()
    xn[3, 0] = float("nan")
    with pytest.raises(AssertionError):           # the reference asserts on NaNs (GAT/layers.py:46-56)
        lay(xn, torch.tensor([3, 1], device=DEV), torch.tensor([1, 2], device=DEV), None)


def test_gat_ode_block_rk4_matches_oracle():
    """GAT-ODE block (8 heads x 16, the config-3 shape), fixed-step rk4, step for step against the restated solver
    driving the oracle function (H independent reference heads, concatenated)."""
    _, _, models = _pkg()
    n, e, d, H = 400, 2500, 128, 8   # 32 GroupNorm groups of 4 channels (1- and 2-channel groups are degenerate)
    oh = d // H
    src, tgt = _random_edges(n, e, seed=3)
    torch.manual_seed(5)
    blk = models.ODEBlock(models.ODEfunc(d, heads=H), method="rk4", options={"step_size": 0.5})
    with torch.no_grad():
        blk.odefunc.norm1.weight.uniform_(0.5, 1.5)
        blk.odefunc.norm1.bias.uniform_(-0.5, 0.5)
    x = torch.randn(n, d)
    gy = torch.randn(n, d)
    pc = {k: v.detach().clone() for k, v in blk.odefunc.state_dict().items()}

    class F_(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ps = torch.nn.ParameterDict({k.replace(".", "_"): torch.nn.Parameter(v) for k, v in pc.items()})
            self.nfe = 0

        def forward(self, t, y):
            self.nfe += 1
            p = self.ps
            yn = torch.nn.functional.group_norm(y, 32, p["norm1_weight"], p["norm1_bias"], 1e-5)
            ttx = torch.cat([torch.ones_like(yn[:, :1]) * t, yn], 1)
            heads = [(p["gc1_f_weight"][h * oh:(h + 1) * oh], p["gc1_f_bias"][h * oh:(h + 1) * oh],
                      p["gc1_w_weight"][h:h + 1], p["gc1_w_bias"][h:h + 1]) for h in range(H)]
            return torch.relu(gat_ref.gat_multihead(ttx, src, tgt, heads))

    fo = F_()
    xo = x.clone().requires_grad_(True)
    yo = oracle_odeint.odeint_adjoint(fo, xo, torch.tensor([0.0, 1.0]), rtol=1e-5, atol=1e-5, method="rk4",
                                      options={"step_size": 0.5})[1]
    yo.backward(gy)
    blk = blk.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    blk.nfe = 0
    y = blk(xg, src.to(DEV), tgt.to(DEV), None)
    nfe_f = blk.nfe
    y.backward(gy.to(DEV))
    assert nfe_f == 8 and blk.nfe == fo.nfe
    G.assert_close(y, yo, rtol=1e-5, atol_scale=1e-5, what="y(1)")
    G.assert_close_l2(xg.grad, xo.grad, 1e-4, what="grad_x")
    for k, p in blk.odefunc.named_parameters():
        if k.endswith("w.bias"):
            continue                                   # mathematically zero (shift invariance of a - max a)
        G.assert_close_l2(p.grad, fo.ps[k.replace(".", "_")].grad, 1e-4, what=k)


@pytest.mark.parametrize("heads", [1, 8])
def test_gat_ode_function_fused_matches_unfused_and_oracle(heads, monkeypatch):
    """The fused ODE-function node (ops.GatOdeFn: no [t || xn] operand, tcgen05 weight gradient from xhat(y)^T dP) against
    the layer-by-layer evaluation (GroupNorm -> cat -> GATconv) and the float64 oracle: value, d/dy, d/dt and every
    parameter gradient.  heads = 1 is the reference's layer (258 projection columns, padded to 260 inside the node)."""
    _, _, models = _pkg()
    n, e, d = 3000, 20000, 128
    oh = d // heads
    src, tgt = _random_edges(n, e, seed=11)
    torch.manual_seed(7)
    f = models.ODEfunc(d, heads=heads)
    with torch.no_grad():
        f.norm1.weight.uniform_(0.5, 1.5)
        f.norm1.bias.uniform_(-0.5, 0.5)
    x = torch.randn(n, d)
    gy = torch.randn(n, d) / n
    p64 = {k: v.detach().double().clone().requires_grad_(True) for k, v in f.state_dict().items()}
    x64 = x.double().requires_grad_(True)
    t64 = torch.tensor(0.37, dtype=torch.float64, requires_grad=True)
    yn = torch.nn.functional.group_norm(x64, 32, p64["norm1.weight"], p64["norm1.bias"], 1e-5)
    hs = [(p64["gc1.f.weight"][h * oh:(h + 1) * oh], p64["gc1.f.bias"][h * oh:(h + 1) * oh], p64["gc1.w.weight"][h:h + 1],
           p64["gc1.w.bias"][h:h + 1]) for h in range(heads)]
    yo = gat_ref.gat_multihead(torch.cat([torch.ones_like(yn[:, :1]) * t64, yn], 1), src, tgt, hs)
    yo.backward(gy.double())
    f = f.to(DEV)
    f.set_adj(src.to(DEV), tgt.to(DEV), None)
    res = {}
    for mode in ("1", "1-gemm", "0"):     # fused with the GroupNorm-fused projection, fused over gn + the general GEMM, layer by layer
        monkeypatch.setenv("GODE_GAT_FUSED", mode[0])
        monkeypatch.setenv("GODE_TC", "6" if mode == "1-gemm" else "7")
        for p in f.parameters():
            p.grad = None
        xg = x.to(DEV).requires_grad_(True)
        tg = torch.tensor(0.37, device=DEV, requires_grad=True)
        y = f(tg, xg)
        y.backward(gy.to(DEV))
        res[mode] = (y.detach(), xg.grad, tg.grad, {k: p.grad.clone() for k, p in f.named_parameters()})
    for mode, (y, gx, gt, gp) in res.items():
        what = {"1": "fused", "1-gemm": "fused (general GEMM)", "0": "unfused"}[mode]
        G.assert_close(y, yo.detach(), rtol=1e-5, atol_scale=1e-5, what=what + " f")
        G.assert_close(gx, x64.grad, rtol=1e-5, atol_scale=1e-5, what=what + " d/dy")
        assert abs(float(gt) - float(t64.grad)) <= 1e-5 * max(abs(float(t64.grad)), 1e-3), (what, float(gt), float(t64.grad))
        for k, v in gp.items():
            if k.endswith("w.bias"):
                continue                               # mathematically zero (shift invariance of a - max a)
            G.assert_close(v, p64[k].grad, rtol=1e-5, atol_scale=1e-5, what=what + " " + k)


def test_gat_model_surface():
    _, layers, models = _pkg()
    m = models.ODEGCN3(nfeat=12, nhid=16, nclass=3, dropout=0.0)
    assert list(m.state_dict())[:4] == ["gc1.f.weight", "gc1.f.bias", "gc1.w.weight", "gc1.w.bias"]
    assert m.gc2.odefunc.gc1.f.weight.shape == (16, 34)
    with pytest.raises(ValueError):
        models.RESK1(12, 16, 3, 0.5, nlayers=2)
    n, e = 300, 1500
    src, tgt = _random_edges(n, e, seed=9)
    m = models.RGCN3(12, 16, 3, 0.0).to(DEV)
    out = m(torch.randn(n, 12, device=DEV), src.to(DEV), tgt.to(DEV), None)
    assert out.shape == (n, 3) and torch.isfinite(out).all()
    out.sum().backward()
    assert all(p.grad is not None for p in m.parameters())
