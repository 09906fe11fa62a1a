"""graph-odenet hot path, B200-native.

Drop-in for the continuous / discrete-residual GNN layer stack of phcavelar/graph-odenet:
``GCN.layers`` / ``GCN.models`` / ``GCN.train_res`` (and ``GAT``, ``QC``) keep the reference's module
surface; the arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of ``libgode.so``
(``include/gode.h``).  There is no CPU fallback: compute calls raise if the library or a CUDA device is
missing.
"""
from . import _lib  # noqa: F401  (loads libgode.so; raises ImportError if it has not been built)

__all__ = ["_lib", "ops", "odeint", "GCN", "GAT", "QC"]
__version__ = "0.1.0"
