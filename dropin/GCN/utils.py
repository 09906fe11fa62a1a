"""Bare-name drop-in for the reference's ``GCN/utils.py`` loaders and metrics (``load_data_new``, ``accuracy``,
``count_params``, ``normalize``, ``sparse_mx_to_torch_sparse_tensor``): bit-exact against the reference's loader
(tests/test_loader_cpu.py) without its matplotlib / removed-scipy imports.  Data is read from ``./data`` like the reference
does (run from the reference's repository root) or from ``$GODE_DATA_ROOT``."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import _root  # noqa: E402,F401
from graph_odenet_b200.utils import (accuracy, count_params, load_data_new, normalize, parse_index_file,  # noqa: E402,F401
                                     sparse_mx_to_torch_sparse_tensor)
