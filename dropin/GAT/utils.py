"""Bare-name drop-in for the reference's ``GAT/utils.py``: ``load_data_new`` returns ``(src, tgt, Mtgt, features, labels,
idx_train, idx_val, idx_test)`` as GAT/utils.py:216 does."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import _root  # noqa: E402,F401
from graph_odenet_b200.utils import accuracy, count_params, normalize, parse_index_file  # noqa: E402,F401
from graph_odenet_b200.utils import load_data_gat as load_data_new  # noqa: E402,F401
