"""Host logic of the solver (graph-odenet_b200/odeint.py) against the restated torchdiffeq (oracle/odeint.py) on CPU: the
generic tuple-state engine runs on plain torch tensors, so step-size selection, accept / reject decisions and NFE can be
compared without a GPU.  ADVICE r01 (low): the initial step of a TUPLE state is 0.01 * max_q(d0_q / d1_q)."""
import math

import numpy as np
import pytest
import torch

import graph_odenet_b200  # noqa: F401
from graph_odenet_b200 import odeint as ours
from oracle import odeint as oracle


class Counter:
    def __init__(self, fn):
        self.fn, self.nfe = fn, 0

    def __call__(self, t, y):
        self.nfe += 1
        return self.fn(t, y)


def _problem(t, y):
    u, v, w = y
    # v starts at 0 with a large derivative (like a_theta in the adjoint): the ratio of maxima would collapse h0
    return (-2.0 * u, 300.0 * torch.cos(5.0 * t) * (1.0 + u.mean()) * torch.ones_like(v), 0.1 * w * torch.tanh(v.mean()))


def _y0():
    g = torch.Generator().manual_seed(0)
    return (torch.randn(40, 8, generator=g), torch.zeros(17), torch.randn(5, generator=g) * 0.1)


@pytest.mark.parametrize("tol", [1e-5, 1e-3])
@pytest.mark.parametrize("t0,t1", [(0.0, 1.0), (1.0, 0.0)])
def test_tuple_state_dopri5_matches_restated_torchdiffeq(tol, t0, t1):
    fo, fu = Counter(_problem), Counter(_problem)
    so, su = {}, {}
    t = torch.tensor([t0, t1])
    want = oracle.odeint(fo, _y0(), t, rtol=tol, atol=tol, method="dopri5", stats=so)
    got = ours._generic_solve(lambda tt, yy: fu(tt, yy), _y0(), t0, t1, "dopri5", None, tol, tol, su)
    assert su == so, (su, so)
    assert fu.nfe == fo.nfe, (fu.nfe, fo.nfe)
    for a, b in zip(got, want):
        # same accept / reject sequence and NFE; the values agree to the solver's GLOBAL error (dt is float32 arithmetic on
        # the host here and on torch scalars there: last-bit differences of dt move the solution inside that error)
        assert torch.allclose(a, b[1], rtol=100 * tol, atol=100 * tol * float(b[1].abs().max()))


def test_initial_h0_semantics():
    h = ours._initial_h0([3.0, 0.0, 2.0], [6.0, 50.0, 1.0])
    assert math.isclose(float(h), 0.02, rel_tol=1e-6)                      # max of the ratios (0.5, 0, 2), not 3 / 50
    assert float(ours._initial_h0([1e-6, 0.0], [1.0, 1.0])) == float(np.float32(1e-6))
    assert math.isinf(float(ours._initial_h0([1.0, 1.0], [1.0, 0.0])))      # x / 0 = inf, as a torch scalar division gives
    assert math.isclose(float(ours._initial_h0([1.0, 0.0], [2.0, 0.0])), 0.005, rel_tol=1e-6)   # trailing 0/0 = nan is ignored by max
