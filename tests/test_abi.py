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
    assert ctypes.sizeof(_lib.SpmmEpilogue) == 8 + 8 + 8 + 8 + 8 * 8 + 4 * 8 + 4 + 4 + 8 + 8 + 8 + 8 + 8 + route
    assert ctypes.sizeof(_lib.Csr) == 2 * 8 + 6 * 8 + 2 * 4
    assert ctypes.sizeof(_lib.GcnOdeFunc) == 2 * ctypes.sizeof(_lib.Csr) + 4 * 4 + 4 * 8 + 8 + 8 + 2 * route
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
