"""Bare-name drop-in for the reference's ``prototypes/orbit/model.py`` (``MLP``, ``IN``, ``IN_ODEfunc``, ``IN_ODE``): what
``prototypes/orbit/train_IN.py:16`` (``from model import IN, IN_ODE``) and ``run_simulation.py`` import."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import _root  # noqa: E402,F401
from graph_odenet_b200.prototypes.orbit.model import MLP, IN, IN_ODEfunc, IN_ODE  # noqa: E402,F401
