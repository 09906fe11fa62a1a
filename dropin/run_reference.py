#!/usr/bin/env python
"""Run one of the reference's OWN, UNMODIFIED scripts on the B200 modules.

    cd /path/to/graph-odenet            # the reference's checkout: its scripts read ./data
    python /path/to/repo/dropin/run_reference.py GCN/train_res.py --model ode3 --dataset cora
    python /path/to/repo/dropin/run_reference.py --keep-models GCN/train_res.py --model ode3   # reference models.py + our layers + solver
    python /path/to/repo/dropin/run_reference.py prototypes/orbit/train_IN.py ...              # ``from model import IN, IN_ODE`` -> dropin/orbit

The reference imports its siblings by bare name (``import models``, ``from utils import ...``, ``from layers import ...``).
Running ``python GCN/train_res.py`` puts the script's directory first on sys.path, so PYTHONPATH cannot override them
(VERDICT r01 weak #12).  This launcher executes the script with ``runpy`` instead, after putting ``dropin/<family>/`` (and
``dropin/`` for the ``torchdiffeq`` stand-in) first on sys.path:

* default: ``layers``, ``models``, ``utils`` all resolve to the shims -> the fused sm_100a engine;
* ``--keep-models``: only ``layers`` / ``utils`` / ``torchdiffeq`` are shimmed and ``models`` stays the reference's file --
  the literal operator-level drop-in (the reference's ODEfunc drives libgode's GraphConvolution kernels through the
  package's generic adjoint engine).
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main(argv):
    keep_models = False
    if argv and argv[0] == "--keep-models":
        keep_models, argv = True, argv[1:]
    if not argv:
        sys.exit(__doc__)
    script = os.path.abspath(argv[0])
    family = os.path.basename(os.path.dirname(script))
    shim_dir = os.path.join(HERE, family)
    if not os.path.isdir(shim_dir):
        sys.exit("no drop-in for %r (have GCN, GAT, QC, orbit)" % family)
    sys.path.insert(0, HERE)                      # torchdiffeq stand-in
    if keep_models:
        # the reference's own models.py, found by bare name AFTER the shims for everything else
        import importlib.util
        sys.path.insert(0, shim_dir)
        spec = importlib.util.spec_from_file_location("models", os.path.join(os.path.dirname(script), "models.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["models"] = mod
        spec.loader.exec_module(mod)
    else:
        sys.path.insert(0, shim_dir)
    # the script's own directory LAST: siblings the drop-in does not replace (prototypes/orbit/prepare_dataset.py ...) still
    # resolve, as they do under ``python script.py``, but never ahead of a shim
    sys.path.append(os.path.dirname(script))
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main(sys.argv[1:])
