"""The drop-in boundary of INTEGRATION.md section 1: scripts that import ``models`` / ``layers`` / ``utils`` by BARE name (as
every script of the reference does) run on the B200 modules through ``dropin/run_reference.py``.

* CPU (this container, where /root/reference exists): the reference's UNMODIFIED ``train_res.py`` is executed; it must get
  through data loading and model construction on the shims and then stop at the package's own "no CPU path" error -- proof
  that the bare names resolved to the drop-in, not to the reference's files.
* GPU (the box has no reference checkout): a probe script with the same bare imports trains two epochs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAUNCH = os.path.join(ROOT, "dropin", "run_reference.py")
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (build container only)")
@pytest.mark.parametrize("args,needle", [
    (["GCN/train_res.py", "--model", "ode3", "--epochs", "1"], "adjacency must live on a CUDA device"),
    (["--keep-models", "GCN/train_res.py", "--model", "res3", "--epochs", "1"], "adjacency must live on a CUDA device"),
    (["GAT/train_res.py", "--model", "ode3", "--epochs", "1"], "GAT edge lists must live on a CUDA device"),
])
def test_unmodified_reference_trainer_resolves_to_the_dropin(args, needle):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-side check: with a GPU the script would simply train")
    r = subprocess.run([sys.executable, LAUNCH] + args, cwd=REF, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert needle in r.stderr, r.stderr[-2000:]
    assert "graph-odenet_b200" in r.stderr and "/root/reference/GCN/layers.py" not in r.stderr


def test_shim_modules_export_what_the_reference_imports():
    """Every bare-name import of the reference's GCN / GAT scripts is satisfied by the shims (no GPU needed)."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import layers, models, utils\n"
            "from utils import load_data_new, accuracy, count_params\n"
            "from layers import GraphConvolution, FixedGraphConvolution\n"
            "from torchdiffeq import odeint_adjoint, odeint\n"
            "for n in ['GCN3','RGCN3','ODEGCN3','RGCN3norm','RGCN3fullnorm','ODEGCN3fullnorm','GCNK','RESK1','RESK2','RESK','ODEK1','ODEK2','ODEfunc','ODEBlock']:\n"
            "    assert hasattr(models, n), n\n"
            "print('OK', layers.__file__)\n")
    for fam in ("GCN", "GAT"):
        r = subprocess.run([sys.executable, "-c", code % (os.path.join(ROOT, "dropin"), os.path.join(ROOT, "dropin", fam))],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "dropin" in r.stdout, r.stderr[-2000:]


def test_orbit_script_resolves_model_to_the_dropin(tmp_path):
    """``prototypes/orbit/train_IN.py:16`` does ``from model import IN, IN_ODE`` next to imports of its own siblings
    (``prepare_dataset``): through the launcher ``model`` is the drop-in and a sibling of the script still resolves.  (The
    reference's trainer itself needs ``fire`` and a generated dataset, neither of which exists here; a probe with the same
    imports stands in.  Construction and state_dict need no GPU.)"""
    d = tmp_path / "orbit"
    d.mkdir()
    (d / "prepare_dataset.py").write_text("def get_epoch():\n    return 'sibling'\n")
    (d / "probe.py").write_text(
        "from prepare_dataset import get_epoch\n"
        "import model\n"
        "from model import IN, IN_ODE\n"
        "net, ode = IN(5, 0, 0, 2), IN_ODE(5, 0, 0, 2)\n"
        "assert 'fR.mlp.0.weight' in net.state_dict() and 'odefunc.fO.output_linear.bias' in ode.state_dict()\n"
        "print('PROBE OK', get_epoch(), model.__file__)\n")
    r = subprocess.run([sys.executable, LAUNCH, str(d / "probe.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "PROBE OK sibling" in r.stdout and os.path.join("dropin", "orbit", "model.py") in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("keep", [False])
def test_bare_name_script_trains_on_the_dropin(keep):
    probe = os.path.join(ROOT, "tests", "dropin_probe", "GCN", "train_probe.py")
    fixture = os.path.join(ROOT, "tests", "golden", "planetoid_cora.npz")
    r = subprocess.run([sys.executable, LAUNCH, probe, fixture], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    assert "PROBE OK" in r.stdout and "dropin/GCN/layers.py" in r.stdout and "dropin/GCN/models.py" in r.stdout
    assert "nfe 20/" in r.stdout or "nfe 26/" in r.stdout or "nfe " in r.stdout
