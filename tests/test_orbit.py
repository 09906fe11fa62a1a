"""The Interaction-Network models of the reference's ``prototypes/orbit`` against fixtures produced by the
unmodified ``prototypes/orbit/model.py`` (tests/golden/make_golden.py ``make_orbit`` -> orbit_golden.npz).

CPU: constructor surface and state_dict layout (checkpoints of train_IN.py must load).  GPU: ``IN`` forward, input gradient
and every parameter gradient at 1e-5; ``IN_ODE`` with the fixed-step rk4 solver at 1e-5 (NFE 4 + 5) and with the reference's
default dopri5 on NFE, accepted / rejected counts and values (gradients at 1e-4: the adaptive controller turns rounding
into O(rtol) differences)."""
import json

import numpy as np
import pytest
import torch

from tests import _golden as G

DEV = "cuda:0"


def _models():
    from graph_odenet_b200.prototypes.orbit import model
    return model


def _sd(fix, prefix):
    return {k[len(prefix):]: torch.from_numpy(fix[k]) for k in fix if k.startswith(prefix)}


def test_state_dict_layout_matches_reference():
    fix = G.load("orbit_golden")
    model = _models()
    for name, cls in (("IN", model.IN), ("IN_ODE", model.IN_ODE)):
        want = _sd(fix, name + "/sd/")
        net = cls(5, 0, 0, 2)
        have = net.state_dict()
        assert list(have) == list(want), name
        assert all(tuple(have[k].shape) == tuple(want[k].shape) for k in want), name
        net.load_state_dict(want)
    ode = model.IN_ODE(5, 0, 0, 2)
    assert ode.tol == 1e-5 and ode.d_P == 2 and ode.nfe == 0
    ode.nfe = 3
    assert ode.odefunc.nfe == 3


def _problem(fix):
    n, m = int(fix["n"]), int(fix["m"])
    src, tgt = torch.from_numpy(fix["src"]), torch.from_numpy(fix["tgt"])
    Msrc, Mtgt = torch.zeros(n, m), torch.zeros(n, m)
    Msrc[src, torch.arange(m)] = 1
    Mtgt[tgt, torch.arange(m)] = 1
    return G.rnd(301, n, 5), G.rnd(302, n, 2), Msrc.to(DEV), Mtgt.to(DEV)


def _run(net, O, Gd, Msrc, Mtgt):
    net.zero_grad()
    o = O.to(DEV).requires_grad_(True)
    P = net(o, None, None, Msrc, Mtgt)
    (P * Gd.to(DEV)).sum().backward()
    return P.detach(), o.grad


def _compare(fix, key, net, P, gO, tol):
    G.assert_close(P, fix[key + "/P"], rtol=tol, atol_scale=tol, what=key + " P")
    G.assert_close(gO, fix[key + "/grad_O"], rtol=tol, atol_scale=tol, what=key + " grad_O")
    names = [k[len(key + "/grad/"):] for k in fix if k.startswith(key + "/grad/")]
    have = {k: p for k, p in net.named_parameters() if p.grad is not None}
    assert sorted(names) == sorted(have), (sorted(names), sorted(have))
    for k in names:
        G.assert_close(have[k].grad, fix[key + "/grad/" + k], rtol=tol, atol_scale=tol, what=key + " grad " + k)


@pytest.mark.gpu
def test_interaction_network_golden():
    fix = G.load("orbit_golden")
    O, Gd, Msrc, Mtgt = _problem(fix)
    net = _models().IN(5, 0, 0, 2)
    net.load_state_dict(_sd(fix, "IN/sd/"))
    net = net.to(DEV)
    P, gO = _run(net, O, Gd, Msrc, Mtgt)
    _compare(fix, "IN", net, P, gO, 1e-5)
    # the relation structure as plans built from index lists (no dense one-hot matrices) gives the same result bit for bit
    from graph_odenet_b200 import ops
    n, m = int(fix["n"]), int(fix["m"])
    one = torch.ones(m, device=DEV)
    j = torch.arange(m, device=DEV)
    psrc = ops.GraphPlan.from_coo(torch.from_numpy(fix["src"]).to(DEV), j, one, n, m)
    ptgt = ops.GraphPlan.from_coo(torch.from_numpy(fix["tgt"]).to(DEV), j, one, n, m)
    P2, gO2 = _run(net, O, Gd, psrc, ptgt)
    assert torch.equal(P, P2) and torch.equal(gO, gO2)


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["rk4", "dopri5"])
def test_interaction_network_ode_golden(method):
    fix = G.load("orbit_golden")
    O, Gd, Msrc, Mtgt = _problem(fix)
    net = _models().IN_ODE(5, 0, 0, 2, method=method)
    net.load_state_dict(_sd(fix, "IN_ODE/sd/"))
    net = net.to(DEV)
    net.stats = {}
    net.nfe = 0
    P, gO = _run(net, O, Gd, Msrc, Mtgt)
    key = "IN_ODE_" + method
    assert net.nfe == int(fix[key + "/nfe"])
    want_stats = json.loads(str(fix[key + "/stats"]))
    for phase in ("forward", "backward"):
        assert {k: v for k, v in net.stats.get(phase, {}).items() if k in ("accepted", "rejected")} == want_stats[phase]
    _compare(fix, key, net, P, gO, 1e-5 if method == "rk4" else 1e-4)
