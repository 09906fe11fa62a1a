"""CPU: the restated Planetoid loader reproduces the reference loader's output (fixtures made by the unmodified
GCN/utils.py:load_data_new and GAT/utils.py:load_data_new, tests/golden/planetoid_<ds>.npz) bit for bit.
Runs only where the reference's data directory exists (the build container); the .npz path is always checked."""
import os

import numpy as np
import pytest
import torch

from tests import _golden as G

REF_DATA = "/root/reference/data"


@pytest.mark.parametrize("ds", ["cora", "citeseer"])
def test_loader_matches_reference_fixture(ds):
    if not os.path.exists(os.path.join(REF_DATA, "ind.%s.allx" % ds)):
        pytest.skip("reference data directory not present")
    from graph_odenet_b200 import utils
    c = G.load("planetoid_" + ds)
    adj, feats, lab, itr, iva, ite = utils.load_data_new(ds, REF_DATA)
    idx, val = adj._indices().numpy(), adj._values().numpy()
    assert np.array_equal(idx[0], c["coo_row"]) and np.array_equal(idx[1], c["coo_col"])       # same (uncoalesced) order
    assert np.array_equal(val.view(np.uint32), c["coo_val"].astype(np.float32).view(np.uint32))  # bit-exact values
    assert np.array_equal(lab.numpy(), c["labels"]) and np.array_equal(ite.numpy(), c["idx_test"])
    assert np.array_equal(itr.numpy(), c["idx_train"]) and np.array_equal(iva.numpy(), c["idx_val"])
    a2, f2, *_ = utils.load_npz(os.path.join(G.HERE, "planetoid_%s.npz" % ds))
    assert torch.equal(f2, feats)
    src, tgt, Mtgt, *_ = utils.load_data_gat(ds, REF_DATA)
    assert np.array_equal(src.numpy(), c["gat_src"]) and np.array_equal(tgt.numpy(), c["gat_tgt"])
    assert Mtgt.shape == (int(c["n"]), len(c["gat_src"]))


def test_npz_loader_shapes():
    from graph_odenet_b200 import utils
    adj, feats, lab, itr, iva, ite = utils.load_npz(os.path.join(G.HERE, "planetoid_cora.npz"))
    assert adj.shape == (2708, 2708) and adj._nnz() == 13264 and feats.shape == (2708, 1433) and int(lab.max()) == 6
    assert len(itr) == 140 and len(iva) == 500 and len(ite) == 1000
    src, tgt, Mtgt, *_ = utils.load_npz(os.path.join(G.HERE, "planetoid_cora.npz"), "GAT")
    assert src.numel() == 5278 and Mtgt.shape == (2708, 5278)
