"""Stand-in for the ``torchdiffeq`` package on top of the package's solvers: ``from torchdiffeq import odeint_adjoint as
odeint`` (GCN/models.py:5) resolves here when ``dropin/`` is on sys.path -- for a reference checkout that keeps its own
``models.py`` and only swaps the operators (``layers.py``) and the solver."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import _root  # noqa: E402,F401
from graph_odenet_b200.odeint import odeint, odeint_adjoint  # noqa: E402,F401
