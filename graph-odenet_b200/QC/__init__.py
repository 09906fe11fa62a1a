"""QC edge-conditioned layer stack with the reference's module surface (QC/layers.py, QC/mpnn.py,
QC/layer_models.py)."""
from . import layers, mpnn, layer_models  # noqa: F401
