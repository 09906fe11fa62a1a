"""The depth sweep's result pickles (GCN/train_layers.py:119-181): record layout, convergence rule, file name
and protocol on the CPU; one tiny sweep end to end on the GPU."""
import os
import pickle
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _tl():
    from graph_odenet_b200 import train_layers
    return train_layers


def test_result_table_layout_and_thresholds():
    tl = _tl()
    rec = tl.new_result_table(3, 6, 4, 20)          # --layers_min 3 --layers_max 5 --runs 4 --epochs 20
    assert sorted(rec) == ["layer_convergence", "layer_test_acc", "layer_test_loss", "layer_val_acc", "layer_val_loss",
                           "max_layers", "min_layers"]
    assert rec["layer_val_acc"].shape == rec["layer_val_loss"].shape == (6, 4, 20)
    assert rec["layer_convergence"].shape == rec["layer_test_acc"].shape == rec["layer_test_loss"].shape == (6, 4)
    assert all(rec[k].dtype == np.float64 for k in rec if k.startswith("layer_"))
    assert (rec["layer_convergence"] == 20).all() and rec["min_layers"] == 3 and rec["max_layers"] == 6
    # GCN/train_layers.py:119-127
    assert tl.thresholds("cora") == (0.7782 * 0.9, 0.7929 * 1.1)
    assert tl.thresholds("citeseer") == (0.6443 * 0.9, 1.2454 * 1.1)
    assert tl.thresholds("pubmed") == (0.7726 * 0.9, 0.7136 * 1.1)
    assert tl.MODEL_NAMES == ["GCNK", "GCNKnorm", "RESK1", "RESK2", "RESK1norm", "RESK2norm", "ODEK1", "ODEK2"]


def test_convergence_rule_freezes_curve_at_previous_epoch(tmp_path):
    tl = _tl()
    rec = tl.new_result_table(3, 5, 2, 6)
    acc_t, loss_t = tl.thresholds("cora")
    curve = [(1.9, 0.2), (1.2, 0.5), (0.8, 0.72), (0.7, 0.9)]
    stopped = None
    for epoch, (loss, acc) in enumerate(curve):
        if tl.record_epoch(rec, 4, 1, epoch, loss, acc, acc_t, loss_t):
            stopped = epoch
            break
    assert stopped == 2 and rec["layer_convergence"][4, 1] == 2
    np.testing.assert_array_equal(rec["layer_val_loss"][4, 1], [1.9, 1.2, 1.2, 1.2, 1.2, 1.2])
    np.testing.assert_array_equal(rec["layer_val_acc"][4, 1], [0.2, 0.5, 0.5, 0.5, 0.5, 0.5])
    assert (rec["layer_val_loss"][3] == 0).all() and rec["layer_convergence"][4, 0] == 6
    # convergence at epoch 0 copies the (still zero) LAST column, as the reference's [epoch - 1] index does
    rec0 = tl.new_result_table(3, 5, 1, 4)
    assert tl.record_epoch(rec0, 3, 0, 0, 0.1, 0.99, acc_t, loss_t)
    np.testing.assert_array_equal(rec0["layer_val_acc"][3, 0], [0, 0, 0, 0])
    # file name and protocol
    path = tl.save_results(rec, "cora", "RESK1", str(tmp_path))
    assert os.path.basename(path) == "cora_RESK1.pickle"
    raw = open(path, "rb").read()
    assert raw[:2] == b"\x80" + bytes([pickle.HIGHEST_PROTOCOL])
    back = tl.load_results(path)
    assert sorted(back) == sorted(rec) and all(np.array_equal(back[k], rec[k]) for k in rec)


@pytest.mark.gpu
def test_tiny_sweep_writes_pickles(tmp_path):
    tl = _tl()
    npz = os.path.join(os.path.dirname(__file__), "golden", "planetoid_cora.npz")
    lines = []
    res = tl.main("GCN", ["--npz", npz, "--runs", "2", "--epochs", "12", "--layers_min", "3", "--layers_max", "4", "--models",
                          "RESK1", "ODEK1", "--out-dir", str(tmp_path)], out=lambda *a, **k: lines.append(" ".join(map(str, a))))
    assert sorted(os.listdir(tmp_path)) == ["cora_ODEK1.pickle", "cora_RESK1.pickle"]
    for m in ("RESK1", "ODEK1"):
        rec = tl.load_results(res["paths"][m])
        assert rec["layer_val_acc"].shape == (5, 2, 12) and rec["max_layers"] == 5
        lo = rec["min_layers"]
        assert 3 <= lo <= 5
        for nl in range(lo, 5):
            assert (rec["layer_test_acc"][nl] > 0.2).all() and np.isfinite(rec["layer_test_loss"][nl]).all()
            assert (rec["layer_val_loss"][nl, :, 0] > 0).all()
        assert (rec["layer_val_acc"][:3] == 0).all()
    assert sum("Test -- epochs:" in ln for ln in lines) >= 4 and sum("Finished!" in ln for ln in lines) == 2
