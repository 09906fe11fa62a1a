"""GPU parity of the row-partitioned path: P ranks == 1 rank on the same graph."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partitioned_plan_world1_matches_plain_plan():
    """One rank, no halo: the partitioned kernel class must reproduce the plain plan bit for bit."""
    import graph_odenet_b200  # noqa: F401
    from graph_odenet_b200 import ops, parallel, synth
    from graph_odenet_b200.GCN import models
    dev = torch.device("cuda:0")
    n, d = 5000, 128
    row, col, val = synth.powerlaw_graph(n, avg_degree=10, seed=2, device=dev)
    torch.manual_seed(0)
    blk = models.ODEBlock(models.ODEfunc(d), method="rk4").to(dev)
    x = torch.randn(n, d, device=dev)
    gy = torch.randn(n, d, device=dev)
    outs = []
    for plan in (ops.GraphPlan.from_coo(row, col, val, n, n), parallel.PartitionedPlan.build(row, col, val, n, 0, 1)):
        for p in blk.parameters():
            p.grad = None
        xx = x.clone().requires_grad_(True)
        y = blk(xx, plan)
        y.backward(gy)
        outs.append((y.detach(), xx.grad, [p.grad.clone() for p in blk.parameters()]))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-5, atol=1e-7)
    for a, b in zip(outs[0][2], outs[1][2]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("method,regime,mode", [("rk4", "relu", "async"), ("rk4", "smooth", "async"), ("rk4", "smooth", "sync"),
                                                ("rk4", "smooth", "split"), ("dopri5", "relu", "async"),
                                                ("dopri5", "smooth", "split"), ("rk4", "smooth", "p2p"),
                                                ("rk4", "relu", "p2p-async"), ("rk4", "smooth", "p2p-async"),
                                                ("dopri5", "smooth", "p2p"), ("dopri5", "relu", "p2p-async"),
                                                ("rk4", "smooth", "p2p-fused"), ("rk4", "relu", "p2p-fused"),
                                                ("dopri5", "smooth", "p2p-fused"),
                                                # forward stages gathered in row chunks, S pushed underneath (GODE_PIPE_G)
                                                ("rk4", "relu", "p2p-fused:g2"), ("rk4", "smooth", "p2p-fused:g3"),
                                                ("dopri5", "relu", "p2p-fused:g2"),
                                                # the supports S exchanged instead of the stage states (GODE_PUSH_Y=0); two-step grid
                                                ("rk4", "relu", "p2p-fused:noY"), ("rk4", "smooth", "p2p-fused:noY"),
                                                ("rk4:0.5", "relu", "p2p-fused"), ("rk4:0.5", "smooth", "p2p-fused:noY"),
                                                ("midpoint", "smooth", "p2p-fused")])
def test_two_gpus_match_one(method, regime, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    world = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "_parallel_worker.py"), method,
           "6000" if method == "dopri5" else "20000", "128", regime]
    env = dict(os.environ, GODE_HALO_MODE=mode.split(":")[0], GODE_PIPE_G=mode.split(":g")[1] if ":g" in mode else "0",
               GODE_PUSH_Y="0" if mode.endswith(":noY") else "1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
