"""Bare-name drop-in for the reference's ``GAT/layers.py`` (same class names, constructors, forward signatures and
state_dict keys); the arithmetic is libgode's sm_100a kernels.  Put this directory FIRST on sys.path (dropin/run_reference.py)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import _root  # noqa: E402,F401
from graph_odenet_b200.GAT.layers import *  # noqa: E402,F401,F403
from graph_odenet_b200.GAT.layers import FixedGraphConvolution, GraphConvolution  # noqa: E402,F401
