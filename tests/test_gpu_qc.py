"""GPU parity of the QC edge-conditioned path (gode_edge_matvec / gode_edge_matvec_bwd + segmented sums behind
graph-odenet_b200/QC) against the golden fixtures produced by the unmodified reference (QC/layers.py) and the
CPU oracle.  Bar: fp32, 1e-5 relative (+1e-5 of the tensor's max magnitude as the absolute floor)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import qc_ref, odeint as oracle_odeint
from tests import _golden as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pkg():
    import graph_odenet_b200  # noqa: F401
    from graph_odenet_b200 import ops, synth
    from graph_odenet_b200.QC import layers, layer_models, mpnn
    return ops, synth, layers, layer_models, mpnn


@pytest.mark.parametrize("nf,nn_,ne", [(24, 60, 130), (73, 18, 16)])
@pytest.mark.parametrize("etgt_form", ["dense", "sparse", "index"])
def test_edgeconv_golden(nf, nn_, ne, etgt_form):
    _, _, layers, _, _ = _pkg()
    g = G.load("qc_golden")
    k = "egc%d/" % nf
    lay = layers.EdgeGraphConvolution(nf, nf).to(DEV)
    lay.load_state_dict(G.params(g, k + "p/"))
    x = G.rnd(70 + nf, nn_, nf).to(DEV).requires_grad_(True)
    ed = G.rnd(90 + nf, ne, nf, nf, scale=1.0 / nf ** 0.5).to(DEV).requires_grad_(True)
    esrc = torch.from_numpy(g[k + "esrc"].astype(np.int64)).to(DEV)
    etgt = torch.from_numpy(g[k + "etgt"].astype(np.int64)).to(DEV)
    if etgt_form == "index":
        Etgt = etgt
    else:
        Etgt = torch.zeros(nn_, ne, device=DEV)                 # the reference's dense one-hot (QC/datasets/utils.py:214)
        Etgt[etgt, torch.arange(ne, device=DEV)] = 1.0
        if etgt_form == "sparse":
            Etgt = Etgt.to_sparse()
    y = lay(x, esrc, Etgt, ed)
    y.backward(G.rnd(95 + nf, nn_, nf).to(DEV))
    G.assert_close(y, g[k + "out"], rtol=1e-5, atol_scale=1e-5, what="out")
    G.assert_close(x.grad, g[k + "grad_x"], rtol=1e-5, atol_scale=1e-5, what="grad_x")
    G.assert_close(ed.grad, g[k + "grad_edge_data"], rtol=1e-5, atol_scale=1e-5, what="grad_edge_data")
    G.assert_close(lay.weight.grad, g[k + "grad_weight"], rtol=1e-5, atol_scale=1e-5, what="grad_weight")
    G.assert_close(lay.bias.grad, g[k + "grad_bias"], rtol=1e-5, atol_scale=1e-5, what="grad_bias")


def test_edge_encoder_golden():
    _, _, layers, _, _ = _pkg()
    g = G.load("qc_golden")
    ee = layers.EdgeEncoderMLP(5, 8).to(DEV)
    ee.load_state_dict(G.params(g, "ee/p/"))        # same state_dict keys as the reference
    out = ee(G.rnd(99, 20, 5).to(DEV))
    G.assert_close(out, g["ee/out"], rtol=1e-5, atol_scale=1e-5, what="edge encoder")


@pytest.mark.parametrize("f", [5, 32, 73, 80, 128])
def test_edge_message_matches_oracle(f):
    """Ragged targets (nodes with 0 and with many incoming edges), any width up to 128, fwd + all gradients."""
    ops = _pkg()[0]
    rs = np.random.RandomState(f)
    n, e = 50, 400
    esrc = torch.from_numpy(rs.randint(0, n, e).astype(np.int64))
    etgt = torch.from_numpy(rs.randint(3, n, e).astype(np.int64))
    etgt[:100] = 7
    s = torch.randn(n, f)
    ed = torch.randn(e, f, f) / f ** 0.5
    b = torch.randn(f)
    gy = torch.randn(n, f)
    so, edo, bo = (t.clone().requires_grad_(True) for t in (s, ed, b))
    m = torch.bmm(edo, so.index_select(0, esrc).unsqueeze(-1)).squeeze(-1)
    yo = torch.zeros(n, f).index_add_(0, etgt, m) + bo
    yo.backward(gy)
    sg, edg, bg = (t.to(DEV).requires_grad_(True) for t in (s, ed, b))
    y = ops.edge_message(sg, edg, esrc.to(DEV), etgt.to(DEV), bg)
    y.backward(gy.to(DEV))
    G.assert_close(y, yo, rtol=1e-5, atol_scale=1e-5, what="out")
    G.assert_close(sg.grad, so.grad, rtol=1e-5, atol_scale=1e-5, what="grad_s")
    G.assert_close(edg.grad, edo.grad, rtol=1e-5, atol_scale=1e-5, what="grad_ed")
    G.assert_close(bg.grad, bo.grad, rtol=1e-5, atol_scale=1e-5, what="grad_b")


def test_edge_message_empty():
    ops = _pkg()[0]
    s = torch.randn(6, 8, device=DEV)
    e0 = torch.empty(0, dtype=torch.int64, device=DEV)
    y = ops.edge_message(s, torch.empty(0, 8, 8, device=DEV), e0, e0, torch.ones(8, device=DEV))
    assert torch.equal(y, torch.ones(6, 8, device=DEV))


def test_edge_gcn_k_sum_matches_oracle():
    ops, synth, _, lm, _ = _pkg()
    hidden = 24
    b = synth.qm9_like_batch(6, hidden, seed=4, device="cpu")
    torch.manual_seed(1)
    model = lm.EdgeGCN_K_Sum(node_features=13, edge_features=5, target_features=12, hidden_features=hidden,
                             num_layers=3, dropout=0.5).eval()
    pc = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    yo = qc_ref.edge_gcn_k_sum(b["node_features"], b["edge_features"], b["esrc"], b["etgt"], b["batch"], pc, 3, hidden)
    tgt = torch.randn_like(yo)
    F.mse_loss(yo, tgt).backward()
    model = model.to(DEV)
    y = model(b["node_features"].to(DEV), b["edge_features"].to(DEV), b["esrc"].to(DEV), b["etgt"].to(DEV),
              b["batch"].to(DEV))
    F.mse_loss(y, tgt.to(DEV)).backward()
    assert y.shape == (6, 12)
    G.assert_close(y, yo, rtol=1e-5, atol_scale=1e-5, what="out")
    for k, p in model.named_parameters():
        G.assert_close(p.grad, pc[k].grad, rtol=1e-4, atol_scale=2e-5, what=k)


def test_scatter_add_known_answer():
    """The only known-answer vectors in the reference: QC/torch_scatter.py:211-218 (scatter_add docstring)."""
    ops = _pkg()[0]
    src = torch.tensor([[2.0, 0, 1, 4, 3], [0, 2, 1, 3, 4]], device=DEV).t().contiguous()   # per-column scatter, dim 0
    index = torch.tensor([4, 5, 4, 2, 3], device=DEV)
    out = ops.scatter_add_rows(src[:, :1].contiguous(), index, 6)
    assert out.squeeze(1).tolist() == [0.0, 0.0, 4.0, 3.0, 3.0, 0.0]


def test_mpnn_enn_edge_matches_restatement():
    ops, synth, _, _, mpnn = _pkg()
    h = 16
    b = synth.qm9_like_batch(4, h, seed=2, device="cpu")
    n, e = b["node_features"].shape[0], b["esrc"].numel()
    torch.manual_seed(3)
    net = mpnn.MPNN_enn_edge(5, h)
    net.set_T(3)
    x = torch.randn(n, h)
    ed = torch.randn(e, h, h) / h ** 0.5
    xo = x
    for _ in range(3):                                   # QC/mpnn.py:26-30
        msg = torch.bmm(ed, xo.index_select(0, b["esrc"]).unsqueeze(-1)).squeeze(-1)
        node_msg = torch.zeros(n, h).index_add_(0, b["etgt"], msg)
        xo = net.update_net(torch.cat([xo, node_msg], 1), xo)
    net = net.to(DEV)
    y = net(x.to(DEV), b["esrc"].to(DEV), b["etgt"].to(DEV), ed.to(DEV))
    G.assert_close(y, xo, rtol=1e-4, atol_scale=1e-5, what="mpnn")


def test_edge_ode_block_matches_oracle():
    """Builder extension (config 5): ODEBlock over the reference EdgeGraphConvolution, rk4, vs the restated solver."""
    ops, synth, _, lm, _ = _pkg()
    d = 128                      # 32 groups of 4 channels (1- and 2-channel groups are degenerate / ill-conditioned)
    b = synth.qm9_like_batch(5, d, seed=6, device="cpu")
    n, e = b["node_features"].shape[0], b["esrc"].numel()
    torch.manual_seed(8)
    blk = lm.EdgeODEBlock(lm.EdgeODEfunc(d), method="rk4")
    ed = torch.randn(e, d, d) / d ** 0.5
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
            yn = F.group_norm(y, lm._groups(d), p["norm1_weight"], p["norm1_bias"], 1e-5)
            ttx = torch.cat([torch.ones_like(yn[:, :1]) * t, yn], 1)
            return F.relu(qc_ref.edge_graph_convolution(ttx, b["esrc"], b["etgt"], ed, p["gc1_weight"], p["gc1_bias"]))

    fo = F_()
    xo = x.clone().requires_grad_(True)
    yo = oracle_odeint.odeint_adjoint(fo, xo, torch.tensor([0.0, 1.0]), rtol=1e-5, atol=1e-5, method="rk4")[1]
    yo.backward(gy)
    blk = blk.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    y = blk(xg, b["esrc"].to(DEV), b["etgt"].to(DEV), ed.to(DEV))
    nfe_f = blk.nfe
    y.backward(gy.to(DEV))
    assert nfe_f == 4 and blk.nfe == fo.nfe == 9
    G.assert_close(y, yo, rtol=1e-5, atol_scale=1e-5, what="y(1)")
    def rel(a, b_):
        return float((a.detach().cpu().double() - b_.double()).norm() / b_.double().norm())

    errs = {"grad_x": rel(xg.grad, xo.grad)}
    errs.update({k: rel(p.grad, fo.ps[k.replace(".", "_")].grad) for k, p in blk.odefunc.named_parameters()})
    print("edge-ODE block gradients, relative L2 vs the restated solver:", errs)
    # measured on B200 (round 2, tcgen05 products with rounded residuals): 2.7e-7 .. 5.1e-7 -- the fp32 bar holds
    G.assert_close_l2(xg.grad, xo.grad, 1e-5, what="grad_x")
    for k, p in blk.odefunc.named_parameters():
        G.assert_close_l2(p.grad, fo.ps[k.replace(".", "_")].grad, 1e-5, what=k)


def test_qc_model_surface():
    _, _, layers, lm, _ = _pkg()
    with pytest.raises(NotImplementedError):
        lm.UnimplementedModel()
    with pytest.raises(ValueError):
        lm.RESKnorm(8, 8, 8, nlayers=2, residue_layers=1)
    with pytest.raises(ValueError):
        lm.RESKnorm(73, 73, 73, nlayers=3)          # GroupNorm(32, 73): the reference cannot build this either
    m = lm.MPNN_ENN_K_Sum(node_features=13, edge_features=5, target_features=12, hidden_features=16)
    assert "mpnn.update_net.weight_ih" in m.state_dict() and "ee.mlp.mlp.layers.0.linear.weight" in m.state_dict()


def test_set2set_golden():
    """Set2Set readout on gode_segment_attend_* against the reference fixture: q*, input gradient, LSTM gradients."""
    _pkg()
    from graph_odenet_b200.QC import set2set
    g = G.load("set2set_golden")
    s2s = set2set.Set2Set(24, 3)
    s2s.load_state_dict(G.params(g, "s2s/p/"))
    s2s = s2s.to(DEV)
    x = torch.from_numpy(g["s2s/x"].copy()).to(DEV).requires_grad_(True)
    batch = torch.from_numpy(g["s2s/batch"].astype(np.int64)).to(DEV)
    out = s2s(x, batch)
    out.backward(torch.from_numpy(g["s2s/g"]).to(DEV))
    G.assert_close(out, g["s2s/out"], rtol=1e-5, atol_scale=1e-5, what="q*")
    G.assert_close(x.grad, g["s2s/grad_x"], rtol=1e-5, atol_scale=1e-5, what="grad_x")
    for k, v in s2s.named_parameters():
        G.assert_close(v.grad, g["s2s/grad/" + k], rtol=1e-5, atol_scale=2e-5, what=k)
    # unsorted batch vector: same result as the sorted one (nodes are regrouped internally)
    perm = torch.randperm(x.shape[0], generator=torch.Generator().manual_seed(1)).to(DEV)
    out2 = s2s(x.detach()[perm], batch[perm])
    G.assert_close(out2, g["s2s/out"], rtol=1e-5, atol_scale=1e-5, what="q* (unsorted batch)")


def test_edge_gcn_k_set2set_golden():
    layer_models = _pkg()[3]
    g = G.load("set2set_golden")
    model = layer_models.EdgeGCN_K_Set2Set(node_features=13, edge_features=5, target_features=12, hidden_features=24,
                                           num_layers=3, s2s_processing_steps=3, type="regression", dropout=0.0)
    model.load_state_dict(G.params(g, "m/p/"))
    model = model.to(DEV).eval()
    batch = torch.from_numpy(g["s2s/batch"].astype(np.int64)).to(DEV)
    y = model(torch.from_numpy(g["m/nf"]).to(DEV), torch.from_numpy(g["m/ef"]).to(DEV),
              torch.from_numpy(g["m/esrc"].astype(np.int64)).to(DEV), torch.from_numpy(g["m/etgt"].astype(np.int64)).to(DEV), batch)
    y.backward(torch.from_numpy(g["m/gy"]).to(DEV))
    G.assert_close(y, g["m/out"], rtol=1e-5, atol_scale=1e-5, what="model out")
    for k, v in model.named_parameters():
        if "m/grad/" + k in g:
            G.assert_close(v.grad, g["m/grad/" + k], rtol=1e-5, atol_scale=5e-5, what=k)


def test_set2set_model_surface():
    _, synth, _, layer_models, _ = _pkg()
    kw = dict(node_features=13, edge_features=5, target_features=12, num_layers=3, s2s_processing_steps=2)
    m = layer_models.MPNN_ENN_K_Set2Set(hidden_features=16, **kw)
    assert {"input.weight", "s2s.lstm.weight_ih_l0", "mpnn.update_net.weight_ih", "output.bias"} <= set(m.state_dict())
    with pytest.raises(ValueError):                       # GroupNorm(32, 73), as in the reference
        layer_models.EdgeRES1_K_Set2Set(hidden_features=73, **kw)
    r = layer_models.EdgeRES1_K_Set2Set(hidden_features=64, **kw).to(DEV).eval()
    b = synth.qm9_like_batch(5, 64, seed=3, device=DEV)
    out = r(b["node_features"], b["edge_features"], b["esrc"], b["etgt"], b["batch"])
    assert out.shape == (5, 12) and bool(torch.isfinite(out).all())
