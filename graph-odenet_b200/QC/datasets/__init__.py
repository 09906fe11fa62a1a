"""Dataset utilities of the QC family (QC/datasets/): the batch collate, on the host as the reference and on the device."""
from . import utils  # noqa: F401
