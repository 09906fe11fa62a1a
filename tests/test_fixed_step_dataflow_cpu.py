"""CPU: the fixed-step solver's data flow (odeint._gcn_fixed_step / _gcn_aug_fixed_step) over a small numeric stand-in for the
fused kernels.  The running final combination (odeint._running_final: the last-but-one stage stores y0 + h*sum b_j k_j instead
of its derivative, the last stage reads that one tensor) and the dropped y(t0) of the adjoint's last step must change nothing but
rounding: in float64 the two forms agree to 1e-13 (state, adjoint state, parameter gradients) and the forward solve equals a
textbook Runge-Kutta step."""
import numpy as np
import pytest
import torch

from graph_odenet_b200 import odeint


class TinyKernel:
    """f(t, y) = tanh(w * y + c * t) per element, with the support S = w * y + c * t as the only tensor between the halves
    (the split the fused kernels use); theta = (w, c) and the time term, all in float64."""

    def __init__(self, n, w, c):
        self.n, self.w, self.c, self.d = n, w, c, 1
        self.dev = torch.device("cpu")
        self.n_theta = 3                       # [d/dw, d/dc, d/dt]
        self.nfe = 0
        self.reads = 0                         # [n]-sized operands read by the fused combinations (the streams)

    def new(self):
        return torch.zeros(self.n, dtype=torch.float64)

    new_S = new_gP = new_Y = new

    def reduce_small(self, t):
        return t

    def transform(self, y, t, out):
        out.copy_(self.w * y + self.c * float(t))
        return out

    def _combine(self, k, y0, kprev, coefs, coef_self, y_next, second):
        if y_next is not None or second is not None:
            self.reads += 1 + len(kprev)
        if y_next is not None:
            y_next.copy_(y0 + sum(float(c) * kp for c, kp in zip(coefs, kprev)) + float(coef_self) * k)
        if second is not None:
            c2, c2s, out2 = second
            assert len(c2) == len(kprev)
            out2.copy_(y0 + sum(float(c) * kp for c, kp in zip(c2, kprev)) + float(c2s) * k)

    def stage_fwd(self, S, k_out, y0=None, kprev=(), coefs=(), coef_self=0.0, y_next=None, t_next=0.0, S_next=None, second=None):
        self.nfe += 1
        k = torch.tanh(S)
        if k_out is not None:
            k_out.copy_(k)
        self._combine(k, y0, kprev, coefs, coef_self, y_next, second)
        if S_next is not None:
            self.transform(y_next, t_next, S_next)

    def vjp_phase1(self, S, a, sign, k_y, gP, y0=None, kprev=(), coefs=(), coef_self=0.0, y_next=None, second=None):
        self.nfe += 1
        k = torch.tanh(S)
        gP.copy_(sign * a * (1 - k * k))
        if k_y is not None:
            k_y.copy_(k)
        self._combine(k, y0, kprev, coefs, coef_self, y_next, second)

    def vjp_phase2(self, y, t, gP, k_a, gtheta, a0=None, kprev=(), coefs=(), coef_self=0.0, a_next=None, second=None):
        ka = gP * self.w
        gtheta[0], gtheta[1], gtheta[2] = (gP * y).sum(), (gP * float(t)).sum(), (gP * self.c).sum()
        if k_a is not None:
            k_a.copy_(ka)
        self._combine(ka, a0, kprev, coefs, coef_self, a_next, second)


def _solve(method, step, running, monkeypatch):
    monkeypatch.setenv("GODE_RK_RUNNING", "1" if running else "0")
    torch.manual_seed(0)
    n = 50
    y0 = torch.randn(n, dtype=torch.float64)
    g1 = torch.randn(n, dtype=torch.float64)
    kf = TinyKernel(n, 0.7, -0.4)
    y1 = odeint.gcn_solve_forward(kf, y0, 0.0, 1.0, method, step)
    kb = TinyKernel(n, 0.7, -0.4)
    a0, ath, at = odeint.gcn_solve_adjoint(kb, y1, g1, 0.0, 1.0, method, step)
    return y0, y1.clone(), a0.clone(), ath.clone(), kf, kb


def _textbook(method, step, y0, w, c):
    tab = odeint.TABLEAUS[method]
    grid = odeint._grid(0.0, 1.0, step)
    y = y0.clone()
    for t0, t1 in zip(grid[:-1], grid[1:]):
        h = float(np.float32(t1 - t0))
        ks = []
        for i in range(tab.s):
            yi = y + sum(h * float(np.float32(tab.a[i][j])) * ks[j] for j in range(i))
            ti = float(np.float32(t0 + np.float32(tab.c[i]) * np.float32(h)))
            ks.append(torch.tanh(w * yi + c * ti))
        y = y + sum(h * float(np.float32(b)) * k for b, k in zip(tab.b, ks))
    return y


@pytest.mark.parametrize("method,step", [("rk4", None), ("rk4", 0.25), ("midpoint", None), ("midpoint", 0.5), ("euler", 0.25)])
def test_running_final_combination_changes_only_rounding(method, step, monkeypatch):
    y0, y1_p, a_p, th_p, kf_p, kb_p = _solve(method, step, False, monkeypatch)
    _, y1_r, a_r, th_r, kf_r, kb_r = _solve(method, step, True, monkeypatch)
    assert kf_p.nfe == kf_r.nfe and kb_p.nfe == kb_r.nfe
    assert float((y1_p - y1_r).abs().max()) < 1e-13
    assert float((a_p - a_r).abs().max()) < 1e-13 and float((th_p - th_r).abs().max()) < 1e-11
    # coefficients are float32 products in the solver (torchdiffeq keeps time in float32): the textbook step with the same
    # float32 step sizes agrees to the rounding of those products
    want = _textbook(method, step, y0, 0.7, -0.4)
    assert float((y1_p - want).abs().max()) < 1e-6
    if method == "rk4":
        # the point of it: fewer operands streamed by the fused combinations (per step: 3 fewer for y, 3 for the adjoint's y --
        # gone altogether on the last step -- and 3 for a)
        assert kf_r.reads < kf_p.reads and kb_r.reads < kb_p.reads
        steps = len(odeint._grid(0.0, 1.0, step)) - 1
        assert kf_p.reads - kf_r.reads == 3 * steps


def test_last_adjoint_step_forms_no_state(monkeypatch):
    """On the last step of the adjoint's grid nothing reads y(t0): phase 1 of the last stage gets no combination at all."""
    monkeypatch.setenv("GODE_RK_RUNNING", "1")
    seen = []

    class Spy(TinyKernel):
        def vjp_phase1(self, S, a, sign, k_y, gP, y0=None, kprev=(), coefs=(), coef_self=0.0, y_next=None, second=None):
            seen.append((y_next is not None, second is not None, k_y is not None))
            return super().vjp_phase1(S, a, sign, k_y, gP, y0, kprev, coefs, coef_self, y_next, second)

    n = 8
    kb = Spy(n, 0.5, 0.1)
    odeint.gcn_solve_adjoint(kb, torch.ones(n, dtype=torch.float64), torch.ones(n, dtype=torch.float64), 0.0, 1.0, "rk4", 0.5)
    # two steps of four stages: first step keeps y (running V at stage 3, y_next at stage 4), the last step drops both
    assert seen[:4] == [(True, False, True), (True, False, True), (True, True, False), (True, False, False)]
    assert seen[4:] == [(True, False, True), (True, False, True), (True, False, False), (False, False, False)]
