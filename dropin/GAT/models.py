"""Bare-name drop-in for the reference's ``GAT/models.py`` (every model class, ``ODEfunc``, ``ODEBlock``)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import _root  # noqa: E402,F401
from graph_odenet_b200.GAT.models import *  # noqa: E402,F401,F403
from graph_odenet_b200.GAT import models as _m  # noqa: E402

globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
