"""GAT layer stack with the reference's module surface (GAT/layers.py, GAT/models.py)."""
from . import layers, models  # noqa: F401
