"""Import shim: ``import graph_odenet_b200`` loads the package that lives in ``graph-odenet_b200/``.

The package directory carries the reference repository's name (with its hyphen), which Python cannot
import directly; this shim registers it under the importable spelling.
"""
import importlib.util as _u
import pathlib as _p
import sys as _s

_real = _p.Path(__file__).resolve().parent.parent / "graph-odenet_b200"
_spec = _u.spec_from_file_location(__name__, _real / "__init__.py", submodule_search_locations=[str(_real)])
_mod = _u.module_from_spec(_spec)
_s.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
