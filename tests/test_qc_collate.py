"""The QC batch collate (QC/datasets/utils.py:153-217) against the reference's own output on synthetic molecules
(tests/golden/qc_collate_golden.npz, written by tests/golden/make_golden.py --only-qc-collate from the unmodified
reference function): the host ``collate_fn`` (CPU test) and the device-side collate of a ragged molecule store
(gode_qc_collate; GPU tests).  Index and copy work: every comparison is bit-exact."""
import numpy as np
import pytest
import torch

from tests import _golden as G

CASES = {"a": {}, "b": {"sizes": (3, 5, 1, 1, 8), "seed": 5}}


def _fixture(tag):
    g = G.load("qc_collate_golden")
    return {k.split("/", 1)[1]: g[k] for k in g if k.startswith(tag + "/")}


def _load_utils():
    import graph_odenet_b200  # noqa: F401  (the C-ABI library loads without a GPU; the host collate calls nothing in it)
    from graph_odenet_b200.QC.datasets import utils
    return utils


@pytest.mark.parametrize("tag", sorted(CASES))
def test_host_collate_matches_reference(tag):
    """``collate_g_concat_edge_data`` (drop-in ``collate_fn``): the reference's 8-tuple, dtype for dtype, bit for bit --
    including the dense block-diagonal G, the dense one-hot E_tgt and the reference's UNSHIFTED edge endpoints."""
    utils = _load_utils()
    want = _fixture(tag)
    bs, G_, B, X, E_d, E_src, E_tgt, Y = utils.collate_g_concat_edge_data(G.synthetic_molecules(**CASES[tag]))
    assert bs == int(want["bs"])
    for name, got in (("G", G_), ("B", B), ("X", X), ("E_d", E_d), ("E_src", E_src), ("E_tgt", E_tgt), ("Y", Y)):
        assert got.numpy().dtype == want[name].dtype and got.shape == want[name].shape, name
        assert np.array_equal(got.numpy(), want[name]), name
    # index-vector form: the row of the single 1 in every column of the dense E_tgt
    _, _, _, _, _, _, idx, _ = utils.collate_g_concat_edge_data(G.synthetic_molecules(**CASES[tag]), dense=False)
    assert np.array_equal(idx.numpy(), want["E_tgt"].argmax(0)) and (want["E_tgt"].sum(0) == 1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(CASES))
def test_device_collate_matches_reference(tag):
    """MoleculeStore.collate (gode_qc_collate) over all molecules in dataset order == the reference collate."""
    import graph_odenet_b200  # noqa: F401
    from graph_odenet_b200.QC.datasets import utils
    want = _fixture(tag)
    mols = G.synthetic_molecules(**CASES[tag])
    store = utils.MoleculeStore(mols, "cuda:0")
    bs, G_, B, X, E_d, E_src, E_tgt, Y = store.collate(list(range(len(mols))))
    assert bs == int(want["bs"]) and G_ is None
    for name, got in (("B", B), ("X", X), ("E_d", E_d), ("E_src", E_src), ("Y", Y)):
        assert np.array_equal(got.cpu().numpy(), want[name]), name
    assert np.array_equal(E_tgt.cpu().numpy(), want["E_tgt"].argmax(0))


@pytest.mark.gpu
def test_device_collate_subsets_and_shift():
    """Arbitrary selections (permuted, repeated, empty) against the host collate of the same list; ``shift=True`` adds the
    molecule's node offset to both endpoints (the block-diagonal batch); the loader covers every molecule once."""
    import graph_odenet_b200  # noqa: F401
    from graph_odenet_b200.QC.datasets import utils
    sizes = tuple(int(v) for v in np.random.RandomState(1).randint(1, 30, 300))
    mols = G.synthetic_molecules(sizes=(4,) + sizes, seed=9)
    store = utils.MoleculeStore(mols, "cuda:0")
    rng = np.random.RandomState(2)
    for ids in (rng.permutation(len(mols))[:57].tolist(), [5, 5, 0, 300, 5], [17], []):
        got = store.collate(ids)
        if not ids:
            assert got[0] == 0 and got[2].numel() == 0 and got[5].numel() == 0
            continue
        want = utils.collate_g_concat_edge_data([mols[i] for i in ids], dense=False)
        assert got[0] == want[0]
        for k in (2, 3, 4, 5, 6, 7):
            assert torch.equal(got[k].cpu(), want[k]), k
        sh = store.collate(ids, shift=True)
        n_of = torch.tensor([mols[i][0][0].shape[0] for i in ids])
        m_of = torch.tensor([len(mols[i][0][2]) for i in ids])
        off = torch.repeat_interleave(torch.cumsum(n_of, 0) - n_of, m_of)
        off2 = torch.cat([off, off])
        assert torch.equal(sh[5].cpu(), want[5] + off2) and torch.equal(sh[6].cpu(), want[6] + off2)
        # with shifted ids every edge stays inside its molecule's block
        assert torch.equal(sh[2][sh[5]].cpu(), sh[2][sh[6]].cpu())
    seen = torch.cat([b[7] for b in utils.DeviceLoader(store, 64, shuffle=True, seed=3)])
    assert seen.shape[0] == len(mols)
    assert torch.equal(torch.sort(seen[:, 0]).values, torch.sort(store.Y_all[:, 0]).values)


@pytest.mark.gpu
def test_device_collate_feeds_the_model():
    """A batch collated on the device runs through EdgeGCN_K_Sum and matches the same batch collated on the host."""
    import graph_odenet_b200  # noqa: F401
    from graph_odenet_b200.QC import layer_models
    from graph_odenet_b200.QC.datasets import utils
    mols = G.synthetic_molecules(sizes=(6, 3, 9, 2, 12, 7), seed=4)
    store = utils.MoleculeStore(mols, "cuda:0")
    torch.manual_seed(0)
    model = layer_models.EdgeGCN_K_Sum(node_features=13, edge_features=5, target_features=12, hidden_features=16, num_layers=2).cuda().eval()
    bs, _, B, X, E_d, E_src, E_tgt, Y = store.collate([3, 0, 5, 1], shift=True)
    out = model(node_features=X, edge_features=E_d, Esrc=E_src, Etgt=E_tgt, batch=B)
    hb = utils.collate_g_concat_edge_data([mols[i] for i in (3, 0, 5, 1)], dense=False)
    n_of = torch.tensor([mols[i][0][0].shape[0] for i in (3, 0, 5, 1)])
    m_of = torch.tensor([len(mols[i][0][2]) for i in (3, 0, 5, 1)])
    off = torch.repeat_interleave(torch.cumsum(n_of, 0) - n_of, m_of)
    off2 = torch.cat([off, off]).cuda()
    out_h = model(node_features=hb[3].cuda(), edge_features=hb[4].cuda(), Esrc=hb[5].cuda() + off2, Etgt=hb[6].cuda() + off2,
                  batch=hb[2].cuda())
    assert out.shape == (4, 12) and torch.equal(out, out_h)
