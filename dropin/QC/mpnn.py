"""Bare-name drop-in for the reference's ``QC/mpnn.py``."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import _root  # noqa: E402,F401
from graph_odenet_b200.QC import mpnn as _m  # noqa: E402

globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
