"""GCN stack of graph-odenet on libgode kernels (mirrors /root/reference/GCN/{layers,models,utils,train_res}.py)."""
from . import layers, models  # noqa: F401
