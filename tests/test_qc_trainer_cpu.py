"""CPU: host logic of the QC experiment driver -- the data-parallel sharding of synthetic molecule batches, the running
average the reference logs with, argument checks."""
import argparse

import pytest
import torch

import graph_odenet_b200  # noqa: F401
from graph_odenet_b200.QC import util


def test_synthetic_shards_partition_the_global_batch():
    full = util.SyntheticLoader(2, 11, 16, seed=5, device="cpu").batches
    for world in (2, 3):
        shards = [util.SyntheticLoader(2, 11, 16, seed=5, device="cpu", rank=r, world=world).batches for r in range(world)]
        for bi, (bs, _, b, x, e_d, e_src, e_tgt, tgt) in enumerate(full):
            assert sum(s[bi][0] for s in shards) == bs == 11
            assert torch.equal(torch.cat([s[bi][3] for s in shards]), x)                 # node features, in order
            assert torch.equal(torch.cat([s[bi][7] for s in shards]), tgt)               # targets, in order
            assert sum(int(s[bi][5].numel()) for s in shards) == int(e_src.numel())      # every edge lives in one shard
            for s in shards:
                n_s, b_s, src_s, tgt_s = s[bi][0], s[bi][2], s[bi][5], s[bi][6]
                assert int(b_s.min()) == 0 and int(b_s.max()) == n_s - 1                  # molecule ids re-based
                assert int(src_s.max()) < b_s.numel() and int(tgt_s.max()) < b_s.numel()
                assert torch.equal(b_s[src_s], b_s[tgt_s])                               # edges stay inside a molecule


def test_average_meter_keeps_the_reference_formula():
    m = util.AverageMeter()
    m.update(1.0, 2)
    m.update(3.0, 2)                 # QC/LogMetric.py:36-38: avg * (count - n) / count + val * n / count, then count += n
    assert (m.val, m.avg, m.count) == (3.0, 3.0, 4)


def test_argument_and_dataset_errors():
    with pytest.raises(argparse.ArgumentTypeError):
        util.restricted_float("0.5", [1e-5, 1e-2])
    assert util.restricted_float("1e-3", [1e-5, 1e-2]) == 1e-3
    for name in ("qm9", "mutag", "enzymes", "other"):
        with pytest.raises(NotImplementedError):
            util.read_dataset(name, "./data", 4, 0)
    with pytest.raises(ValueError):
        util.get_metric_by_task_type("ranking", 1)
    crit, ev, name, better, best = util.get_metric_by_task_type("regression", 12)
    assert name == "Error ratio" and better(1.0, 2.0) and best(1.0, 2.0) == 1.0
