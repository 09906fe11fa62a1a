"""CPU: the C-ABI library loads and exports exactly what include/gode.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "graph-odenet_b200", "csrc", "libgode.so")


def _declared():
    src = open(os.path.join(ROOT, "include", "gode.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gode_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(LIB)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "include/gode.h declares %s but libgode.so does not export it" % n


def test_binding_covers_header():
    from graph_odenet_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()
    assert _lib.lib.gode_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA tensor the product path must fail loudly, not compute on the CPU."""
    import torch
    from graph_odenet_b200 import ops
    from graph_odenet_b200.GCN import layers
    with pytest.raises(TypeError):
        ops.gemm(torch.ones(2, 2), torch.ones(2, 2))
    lay = layers.GraphConvolution(4, 4)
    adj = torch.eye(4).to_sparse()
    with pytest.raises(TypeError):
        lay(torch.ones(4, 4), adj)


def test_struct_layouts_match_header():
    """ctypes mirrors of the two ABI structs have the C layout (sizes computed from the header's field order)."""
    from graph_odenet_b200 import _lib
    route = 8 + 8 + 16 * 8
    assert ctypes.sizeof(_lib.PushRoute) == route
    second = 4 * 8 + 4 + 4 + 8        # coef[8], coef_self, padding, out
    assert ctypes.sizeof(_lib.RkSecond) == second
    assert ctypes.sizeof(_lib.SpmmEpilogue) == 8 + 8 + 8 + 8 + 8 * 8 + 4 * 8 + 4 + 4 + 8 + 8 + 8 + 8 + 8 + 2 * route + second + 8
    assert ctypes.sizeof(_lib.Csr) == 2 * 8 + 6 * 8 + 2 * 4 + 8 + 4 + 4     # ..., tile_sched, n_tile_sched, padding
    assert ctypes.sizeof(_lib.GcnOdeFunc) == 2 * ctypes.sizeof(_lib.Csr) + 4 * 4 + 4 * 8 + 8 + 8 + 3 * route + second + 8
    assert ctypes.sizeof(_lib.PeerGroup) == 4 + 4 + 16 * 8


def test_model_surface_matches_reference_keys():
    from graph_odenet_b200.GCN import models
    m = models.ODEGCN3(nfeat=10, nhid=16, nclass=3, dropout=0.5)
    assert list(m.state_dict()) == ["gc1.weight", "gc1.bias", "gc2.odefunc.norm1.weight", "gc2.odefunc.norm1.bias",
                                    "gc2.odefunc.gc1.weight", "gc2.odefunc.gc1.bias", "gc3.weight", "gc3.bias"]
    assert m.gc2.odefunc.gc1.weight.shape == (17, 16) and m.nfe == 0
    m.nfe = 3
    assert m.gc2.odefunc.nfe == 3
    for cls, kw in ((models.GCNK, dict(nlayers=1)), (models.RESK1, dict(nlayers=2)), (models.RESK2, dict(nlayers=3)),
                    (models.RESK, dict(nlayers=3, residue_layers=2)), (models.ODEK1, dict(nlayers=2)),
                    (models.ODEK2, dict(nlayers=3))):
        with pytest.raises(ValueError):
            cls(10, 16, 3, 0.5, **kw)
    with pytest.raises(ValueError):
        models.RGCN2(10, 2, 3, 0.5)
    k = models.RESKnorm(10, 16, 3, 0.5, nlayers=5, residue_layers=3)
    assert [n for n, _ in k.named_parameters()][:2] == ["gcs.0.weight", "gcs.0.bias"] and len(k.norms) == 3


def test_tile_schedule_is_a_merge_of_tiles_and_hub_chunk_groups():
    """gode_csr_t.tile_sched (ops.tile_schedule): every 32-row tile once, in id order; every group of 8 hub chunks once, right
    behind the tile that contains the hub of its first chunk (index work, checked on CPU tensors)."""
    import torch
    from graph_odenet_b200 import ops
    n_rows = 10_000
    heavy = torch.tensor([5, 31, 32, 700, 701, 9_999], dtype=torch.int32)
    nch = torch.tensor([3, 1, 17, 2, 40, 9])
    chunk_ptr = torch.cat([torch.zeros(1, dtype=torch.int64), nch.cumsum(0)]).to(torch.int32)
    n_chunks = int(nch.sum())
    sched = ops.tile_schedule(n_rows, heavy, chunk_ptr, len(heavy), n_chunks)
    n_tiles, groups = (n_rows + 31) // 32, (n_chunks + 7) // 8
    assert sched.dtype == torch.int32 and sched.numel() == n_tiles + groups
    tiles = sched[sched >= 0]
    assert torch.equal(tiles, torch.arange(n_tiles, dtype=torch.int32))
    grp = (-sched[sched < 0] - 1).tolist()
    assert sorted(grp) == list(range(groups))
    pos = {int(v): i for i, v in enumerate(sched.tolist())}
    for g in range(groups):
        hub = int(torch.searchsorted(chunk_ptr.to(torch.int64), torch.tensor([8 * g]), right=True)) - 1
        tile = int(heavy[hub]) // 32
        assert pos[tile] < pos[-(g + 1)] and (tile + 1 == n_tiles or pos[-(g + 1)] < pos[tile + 1])
    assert ops.tile_schedule(n_rows, None, None, 0, 0) is None


def test_row_stochastic_detection_decides_the_pattern_only_transpose():
    """ops.GraphPlan.row_stochastic_scale: the adjoint may gather A_hat^T gP on the 0/1 pattern (and take the bias gradient from
    the column sums of the result) only when every row of A_hat is constant AND sums to one -- an empty row, or a scaled matrix,
    must keep the values (index / float work on CPU tensors through an uninitialised plan object)."""
    import numpy as np
    import torch
    from graph_odenet_b200 import ops

    def plan(rowptr, row_vals):
        p = object.__new__(ops.GraphPlan)
        p.n_rows = len(rowptr) - 1
        p.rowptr = torch.tensor(rowptr, dtype=torch.int32)
        p.row_vals = None if row_vals is None else torch.tensor(row_vals, dtype=torch.float32)
        return p

    third = float(np.float32(1.0 / 3.0))
    ok = plan([0, 2, 3, 6], [0.5, 1.0, third])
    assert ok.row_stochastic_scale() is ok.row_vals
    assert plan([0, 2, 3, 6], [0.5, 1.0, 0.25]).row_stochastic_scale() is None          # a row that sums to 0.75
    assert plan([0, 2, 2, 5], [0.5, 0.0, third]).row_stochastic_scale() is None         # an empty row
    assert plan([0, 2, 3, 6], [0.25, 0.5, third / 2]).row_stochastic_scale() is None    # 0.5 * A_hat
    assert plan([0, 2, 3, 6], None).row_stochastic_scale() is None                      # not row-constant
    # degrees up to 100 000: fp32(1 / deg) * deg stays within the 1e-6 window
    deg = torch.tensor([1, 2, 3, 7, 1000, 99_991, 100_000])
    big = object.__new__(ops.GraphPlan)
    big.n_rows = deg.numel()
    big.rowptr = torch.cat([torch.zeros(1, dtype=torch.int64), deg.cumsum(0)]).to(torch.int32)
    big.row_vals = (1.0 / deg.double()).float()
    assert big.row_stochastic_scale() is big.row_vals
