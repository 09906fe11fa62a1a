"""ODE solvers of the hot path: ``odeint`` / ``odeint_adjoint`` with torchdiffeq's call signature.

replaces: ``from torchdiffeq import odeint_adjoint as odeint`` (GCN/models.py:5) and the call
``odeint(self.odefunc, x, self.integration_time, rtol=self.tol, atol=self.tol)`` (GCN/models.py:192).

Two engines share the Butcher tableaux and the step-size controller below:

* **fused GCN engine** -- when ``func`` is the GCN ``ODEfunc`` (GCN/models.py:161-179) every function
  evaluation is two libgode calls (``gode_gcn_transform`` + ``gode_gcn_stage_fwd``; the adjoint adds
  ``gode_gcn_vjp_phase2``) and the Runge-Kutta stage combination ``y0 + dt*sum a_ij k_j`` is computed in the
  epilogue of the SpMM that produces ``k_i`` -- stage tensors are never combined by separate elementwise ops;
* **generic engine** -- any other ``nn.Module`` ODE function (``ODEfunc2``, the GAT and QC functions): tuple
  states, the module's own forward (whose layers are libgode kernels) and ``torch.autograd.grad`` for the
  adjoint, with libgode's ``gode_rk_combine`` / ``gode_rk_error_sumsq`` doing the [N,d] stage arithmetic.

Methods: ``euler``, ``midpoint``, ``rk4`` (torchdiffeq's 3/8-rule step), ``dopri5`` (default, adaptive).
Time values and step sizes are float32 on the host, as torchdiffeq keeps them in float32 tensors.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os

import numpy as np
import torch

from . import _lib, ops
from ._lib import lib, check

F32 = np.float32

# Test instrumentation (tests/test_gpu_gcn.py::test_mask_conditioned_parity): when set to a list, the fused GCN engine
# appends the ReLU mask (k > 0) of every function evaluation to it, in evaluation order, so that the CPU oracle can be run
# on exactly the masks the CUDA path used (the parity protocol SURVEY 8c(5) applies to dropout masks).
MASK_LOG = None

# ---------------------------------------------------------------------------------------------
# Butcher tableaux: c (nodes), a (strictly lower rows), b (weights)
# ---------------------------------------------------------------------------------------------


class Tableau:
    def __init__(self, c, a, b, c_err=None, c_mid=None, fsal=False, order=1):
        self.c, self.a, self.b, self.c_err, self.c_mid, self.fsal, self.order = c, a, b, c_err, c_mid, fsal, order
        self.s = len(b)


_DP_A = [
    [],
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]

TABLEAUS = {
    "euler": Tableau([0.0], [[]], [1.0], order=1),
    "midpoint": Tableau([0.0, 0.5], [[], [0.5]], [0.0, 1.0], order=2),
    # 3/8 rule -- what torchdiffeq's method='rk4' steps with
    "rk4": Tableau([0.0, 1 / 3, 2 / 3, 1.0], [[], [1 / 3], [-1 / 3, 1.0], [1.0, -1.0, 1.0]],
                   [1 / 8, 3 / 8, 3 / 8, 1 / 8], order=4),
    "dopri5": Tableau(
        [0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0], _DP_A,
        [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0],
        c_err=[35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
               -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0],
        c_mid=[6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
               187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2],
        fsal=True, order=5),
}
FIXED_METHODS = ("euler", "midpoint", "rk4")


def _grid(t0, t1, step_size):
    """Fixed-step time grid (float32), torchdiffeq's constructor: arange(t0, t1, h) with the end clamped."""
    t0, t1 = F32(t0), F32(t1)
    if step_size is None:
        return [t0, t1]
    h = F32(step_size)
    sign = F32(1.0) if t1 >= t0 else F32(-1.0)
    n = int(np.ceil(F32(abs(t1 - t0)) / h + F32(1.0)))
    g = [F32(t0 + sign * F32(i) * h) for i in range(n)]
    if (sign > 0 and g[-1] > t1) or (sign < 0 and g[-1] < t1):
        g[-1] = t1
    return g


def _optimal_step(last, msr, order=5, safety=0.9, ifactor=10.0, dfactor=0.2):
    msr = F32(msr)
    if msr == 0:
        return F32(last * F32(ifactor))
    if msr < 1:
        dfactor = 1.0
    ratio = F32(np.sqrt(msr))
    factor = max(F32(1.0 / ifactor), min(F32(ratio ** F32(1.0 / order)) / F32(safety), F32(1.0 / dfactor)))
    return F32(last / factor)


def _initial_h0(d0s, d1s):
    """Hairer's first guess for a TUPLE state as torchdiffeq 0.0.x forms it: ``1e-6`` if the largest scaled norm of the
    state or of its derivative is below 1e-5, else ``0.01 * max_q(d0_q / d1_q)`` -- the maximum of the per-tensor RATIOS,
    not the ratio of the maxima (in the adjoint a_theta(t1) = 0 while its derivative is large; the ratio of maxima would
    let that tensor collapse h0).  Division and ``max`` follow IEEE / Python semantics as torch scalars do there: x/0 is
    inf, 0/0 is nan, and a nan is kept only if it comes first."""
    if max(d0s) < 1e-5 or max(d1s) < 1e-5:
        return F32(1e-6)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratios = [float(np.float64(a) / np.float64(b)) for a, b in zip(d0s, d1s)]
    return F32(0.01 * max(ratios))


def _interp_weights(x, dt, c_mid):
    """Quartic dense output of dopri5 written as weights on (y0, y1, k_0..k_6)."""
    x = float(x)
    x2, x3, x4 = x * x, x ** 3, x ** 4
    w_y0 = -8 * x4 + 18 * x3 - 11 * x2 + 1
    w_y1 = -8 * x4 + 14 * x3 - 5 * x2
    w_ym = 16 * x4 - 32 * x3 + 16 * x2
    w_f0 = float(dt) * (-2 * x4 + 5 * x3 - 4 * x2 + x)
    w_f1 = float(dt) * (2 * x4 - 3 * x3 + x2)
    wk = [w_ym * float(dt) * cm for cm in c_mid]
    wk[0] += w_f0
    wk[-1] += w_f1
    return w_y0 + w_ym, w_y1, wk


def _times(t):
    if torch.is_tensor(t):
        return [F32(v) for v in t.detach().cpu().tolist()]
    return [F32(v) for v in t]


# ---------------------------------------------------------------------------------------------
# fused GCN engine
# ---------------------------------------------------------------------------------------------


class GcnKernel:
    """Binds a GCN ODE function's parameters and graph plan to the ``gode_gcn_*`` entry points."""

    def __init__(self, plan, weight, bias, gamma, beta, groups, eps=1e-5, precision=_lib.PREC_FP32):
        self.plan = plan
        self.d = int(weight.shape[1])
        if weight.shape[0] != self.d + 1:
            raise ValueError("ODE function weight must be [d+1, d]")
        self.n = plan.n_rows
        self.dev = weight.device
        self.weight, self.bias, self.gamma, self.beta = (weight.detach().contiguous(),
                                                         None if bias is None else bias.detach().contiguous(),
                                                         gamma.detach().contiguous(), beta.detach().contiguous())
        f = _lib.GcnOdeFunc()
        f.A = plan.csr(False)
        if plan.rowptr_t is not None:
            f.At = plan.csr(True)
        f.d, f.groups, f.gn_eps, f.precision = self.d, groups, eps, precision
        f.W, f.gamma, f.beta = self.weight.data_ptr(), self.gamma.data_ptr(), self.beta.data_ptr()
        f.b = self.bias.data_ptr() if self.bias is not None else None
        # Row-stochastic, row-constant A_hat (the reference's normalisation): the adjoint's A_hat^T gather runs on the 0/1
        # pattern with gP pre-scaled by its producer, and the bias gradient comes out of the weight-gradient pass
        # (gode_gcn_odefunc_t.gp_row_scale; GODE_UNIT_T=0 keeps values per entry and the separate column-sum pass).
        self.unit_t = None
        if plan.rowptr_t is not None and os.environ.get("GODE_UNIT_T", "1") != "0" and hasattr(plan, "unit_transpose"):
            self.unit_t = plan.unit_transpose()
            if self.unit_t is not None:
                f.At = self.unit_t[1]
                f.gp_row_scale = self.unit_t[0].data_ptr()
        self.f = f
        self.ws_bytes = lib.gode_gcn_workspace_bytes(C.byref(f))
        self.n_theta = (self.d + 1) * self.d + 3 * self.d + 1
        self.nfe = 0

    def _ws(self):
        return ops.workspace(self.ws_bytes, self.dev, "gcn")

    def new(self):
        return torch.empty(self.n, self.d, dtype=torch.float32, device=self.dev)

    # Hooks the row-partitioned kernel (parallel.PartitionedGcnKernel) overrides: operand buffers of the two
    # gathers carry a halo tail there, and reductions over all nodes / all ranks need a collective.
    def new_S(self):
        """Buffer for a support S (operand of the A_hat gather)."""
        return self.new()

    def new_gP(self):
        """Buffer for gP (operand of the A_hat^T gather)."""
        return self.new()

    def new_Y(self):
        """Buffer for a stage state the NEXT evaluation's transform reads (the row-partitioned kernel hands out one with a
        halo tail: the producing gather stores the peers' rows there, parallel.HaloKernelMixin)."""
        return self.new()

    def transform_rows(self, y, t, out, row0, n_rows):
        """out[row0 : row0 + n_rows] = transform(y[row0 : row0 + n_rows], t)  (base pointers; rows past the owned block
        are the halo rows of a partitioned block)."""
        ws = self._ws()
        check(lib.gode_gcn_transform_rows(C.byref(self.f), ops._p(y), float(t), ops._p(out), int(row0), int(n_rows),
                                          ops._p(ws), self.ws_bytes, ops._stream()), "gode_gcn_transform_rows")
        return out

    def numel_global(self):
        """Number of state elements over the whole graph (denominator of the RMS norms)."""
        return self.n * self.d

    def scalar(self, dev_scalar):
        """Host value of a device scalar that is a sum over nodes."""
        return float(dev_scalar.item())

    def reduce_small(self, t):
        """In-place sum over ranks of a small tensor of per-rank partial sums (a_theta, a_t)."""
        return t

    def transform(self, y, t, out):
        ws = self._ws()
        check(lib.gode_gcn_transform(C.byref(self.f), ops._p(y), float(t), ops._p(out), ops._p(ws), self.ws_bytes,
                                     ops._stream()), "gode_gcn_transform")
        return out

    @contextlib.contextmanager
    def _second(self, second, n_prev):
        """The descriptor's per-call second combination ``(coefs2, coef2_self, out2)`` -- out2 = y0 + sum coefs2*kprev +
        coef2_self*k over the SAME (y0, kprev) as the call's own combination -- set for the one library call inside the block."""
        if second is None:
            yield
            return
        c2, c2s, out2 = second
        if len(c2) != n_prev:
            raise ValueError("second combination: %d coefficients for %d stage tensors" % (len(c2), n_prev))
        r = _lib.RkSecond()
        for j, c in enumerate(c2):
            r.coef[j] = float(c)
        r.coef_self, r.out = float(c2s), out2.data_ptr()
        self.f.second = r
        try:
            yield
        finally:
            self.f.second = _lib.RkSecond()

    def stage_fwd(self, S, k_out, y0=None, kprev=(), coefs=(), coef_self=0.0, y_next=None, t_next=0.0, S_next=None,
                  second=None):
        """k_out = f(.) from its support S; fused y_next = y0 + sum coefs*kprev + coef_self*k; S_next = transform;
        ``second`` = (coefs2, coef2_self, out2): a second combination of the same operands (see ``_second``)."""
        self.nfe += 1
        ws = self._ws()
        if MASK_LOG is not None and k_out is None:
            k_out = self.new()
        karr = (C.c_void_p * _lib.MAX_STAGES)(*[k.data_ptr() for k in kprev])
        carr = (C.c_float * _lib.MAX_STAGES)(*[float(c) for c in coefs])
        with self._second(second, len(kprev)):
            check(lib.gode_gcn_stage_fwd(C.byref(self.f), ops._p(S), ops._p(k_out), ops._p(y0), karr, carr, len(kprev),
                                         float(coef_self), ops._p(y_next), float(t_next), ops._p(S_next), ops._p(ws),
                                         self.ws_bytes, ops._stream()), "gode_gcn_stage_fwd")
        if MASK_LOG is not None:
            MASK_LOG.append(k_out > 0)

    def vjp_phase1(self, S, a, sign, k_y, gP, y0=None, kprev=(), coefs=(), coef_self=0.0, y_next=None, second=None):
        """k_y = f(.), gP = sign*a*(k_y>0), fused y_next = y0 + sum coefs*kprev + coef_self*k_y -- one SpMM launch."""
        self.nfe += 1
        ws = self._ws()
        if MASK_LOG is not None and k_y is None:
            k_y = self.new()
        karr = (C.c_void_p * _lib.MAX_STAGES)(*[k.data_ptr() for k in kprev])
        carr = (C.c_float * _lib.MAX_STAGES)(*[float(c) for c in coefs])
        with self._second(second, len(kprev)):
            check(lib.gode_gcn_vjp_phase1(C.byref(self.f), ops._p(S), ops._p(a), float(sign), ops._p(k_y), ops._p(gP),
                                          ops._p(y0), karr, carr, len(kprev), float(coef_self), ops._p(y_next),
                                          ops._p(ws), self.ws_bytes, ops._stream()), "gode_gcn_vjp_phase1")
        if MASK_LOG is not None:
            MASK_LOG.append(k_y > 0)

    def vjp_phase2(self, y, t, gP, k_a, gtheta, a0=None, kprev=(), coefs=(), coef_self=0.0, a_next=None, second=None):
        """k_a = (gP^T A_hat) d[.]/dy (k_a may be None when no later stage reads it), gtheta = parameter / time terms;
        fused a_next = a0 + sum coefs*kprev + coef_self*k_a in the tail of the GroupNorm backward (``second``: as in
        ``stage_fwd``, over a0 / kprev / k_a)."""
        ws = self._ws()
        karr = (C.c_void_p * _lib.MAX_STAGES)(*[k.data_ptr() for k in kprev])
        carr = (C.c_float * _lib.MAX_STAGES)(*[float(c) for c in coefs])
        with self._second(second, len(kprev)):
            check(lib.gode_gcn_vjp_phase2_rk(C.byref(self.f), ops._p(y), float(t), ops._p(gP), ops._p(k_a), ops._p(gtheta),
                                             ops._p(a0), karr, carr, len(kprev), float(coef_self), ops._p(a_next),
                                             ops._p(ws), self.ws_bytes, ops._stream()), "gode_gcn_vjp_phase2_rk")


def _nz(ks, coefs):
    """Drop zero coefficients (and the tensors they multiply)."""
    kk, cc = [], []
    for k, c in zip(ks, coefs):
        if c != 0:
            kk.append(k)
            cc.append(c)
    return kk, cc


def _needed_later(tab, j):
    """Is k_j read by any later stage combination or by the final weights?"""
    return any(len(tab.a[i]) > j and tab.a[i][j] != 0 for i in range(j + 1, tab.s)) or tab.b[j] != 0 or (
        tab.c_err is not None and tab.c_err[j] != 0) or (tab.c_mid is not None and tab.c_mid[j] != 0)


def gcn_solve_forward(kern, y0, t0, t1, method="dopri5", step_size=None, rtol=1e-5, atol=1e-5, stats=None):
    """Integrate y' = f(t, y) from t0 to t1 on the fused kernels; returns y(t1)."""
    tab = TABLEAUS[method]
    if method in FIXED_METHODS:
        grid = _grid(t0, t1, step_size)
        y = y0
        S = kern.transform(y, grid[0], kern.new_S())
        for g0, g1 in zip(grid[:-1], grid[1:]):
            dt = F32(g1 - g0)
            last_step = g1 == grid[-1]
            y, S = _gcn_fixed_step(kern, tab, g0, dt, y, S, want_S=not last_step)
            if stats is not None:
                stats["accepted"] = stats.get("accepted", 0) + 1
        return y
    return _gcn_dopri5(kern, tab, y0, F32(t0), F32(t1), rtol, atol, stats)


def _running_final(tab):
    """May the step keep a running partial sum of its final combination instead of the last-but-one stage derivative?
    In an explicit s-stage step k_{s-2} is read, after its own stage's combination, by the final weights b only; so that
    stage stores  V = y0 + dt * sum_{j <= s-2} b_j k_j  in place of k_{s-2} (a second combination of operands it reads
    anyway) and the last stage computes y1 = V + dt * b_{s-1} k_{s-1} from ONE tensor instead of y0 and every k_j -- for
    rk4 three [N, d] streams less per state and step.  GODE_RK_RUNNING=0 restores the plain form."""
    return (tab.s >= 2 and any(c != 0 for c in tab.b[:tab.s - 1]) and tab.c_err is None
            and os.environ.get("GODE_RK_RUNNING", "1") != "0")


def _stage_plan(tab, i, dt, ks, running):
    """Operands of stage i's fused combinations: (base is y0?, kprev, coefs, coef_self, second coefficients or None)."""
    s = tab.s
    last = i == s - 1
    row = tab.b if last else tab.a[i + 1]
    coefs = [F32(dt * F32(c)) for c in row[:i + 1]]
    if running and last:
        # y1 = V + dt * b_{s-1} * k_{s-1}
        return False, [], [], coefs[i], None
    if running and i == s - 2:
        b = [F32(dt * F32(c)) for c in tab.b[:i + 1]]
        keep = [j for j in range(i) if coefs[j] != 0 or b[j] != 0]
        return True, [ks[j] for j in keep], [coefs[j] for j in keep], coefs[i], ([b[j] for j in keep], b[i])
    kprev, cprev = _nz(ks[:i], coefs[:i])
    return True, kprev, cprev, coefs[i], None


def _gcn_fixed_step(kern, tab, t0, dt, y0, S0, want_S):
    s = tab.s
    ks, S = [], S0
    scratch_y = [kern.new_Y(), kern.new_Y()]
    y1 = V = None
    running = _running_final(tab)
    for i in range(s):
        last = i == s - 1
        t_n = F32(t0 + dt) if last else F32(t0 + F32(tab.c[i + 1]) * dt)
        base_y0, kprev, cprev, c_self, sec = _stage_plan(tab, i, dt, ks, running)
        store = (not last) and _needed_later(tab, i) and sec is None
        k_i = kern.new() if store else None
        y_next = (kern.new_Y() if want_S else kern.new()) if last else scratch_y[i & 1]
        S_next = kern.new_S() if (not last or want_S) else None
        second = None
        if sec is not None:
            V = kern.new()
            second = (sec[0], sec[1], V)
        kern.stage_fwd(S, k_i, y0 if base_y0 else V, kprev, cprev, c_self, y_next, t_n, S_next, second=second)
        ks.append(k_i)
        S = S_next
        y1 = y_next
    return y1, S


def _msr(kern, y0, y1, ks, coefs, rtol, atol):
    kk, cc = _nz(ks, coefs)
    return ops.rk_error_sumsq(y0, y1, kk, cc, rtol, atol)


def _gcn_dopri5(kern, tab, y0, t0, t1, rtol, atol, stats):
    n_el = kern.numel_global()
    S = kern.transform(y0, t0, kern.new_S())
    f0 = kern.new()
    kern.stage_fwd(S, f0)
    # Hairer initial step: norms are RMS of x / (atol + rtol*|y0|)
    d0 = float(np.sqrt(kern.scalar(ops.rk_error_sumsq(y0, y0, [y0], [1.0], rtol, atol)) / n_el))
    d1 = float(np.sqrt(kern.scalar(ops.rk_error_sumsq(y0, y0, [f0], [1.0], rtol, atol)) / n_el))
    h0 = F32(1e-6) if (d0 < 1e-5 or d1 < 1e-5) else F32(0.01 * d0 / d1)
    y_probe, S_probe, f_probe = kern.new(), kern.new_S(), kern.new()
    ops.rk_combine(y0, [f0], [h0], out=y_probe)
    kern.transform(y_probe, F32(t0 + h0), S_probe)
    kern.stage_fwd(S_probe, f_probe)
    d2 = float(np.sqrt(kern.scalar(ops.rk_error_sumsq(y0, y0, [f_probe, f0], [1.0, -1.0], rtol, atol)) / n_el)) / float(h0)
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(F32(1e-6), F32(h0 * F32(1e-3)))
    else:
        h1 = F32((0.01 / max(d1, d2)) ** (1.0 / 5.0))
    dt = F32(min(F32(100.0) * h0, h1))
    del y_probe, S_probe, f_probe

    t = t0
    y, f = y0, f0
    ks = [f] + [kern.new() for _ in range(6)]
    Yb, Sb = [kern.new(), kern.new()], [kern.new_S(), kern.new_S()]
    last = None  # (t_lo, t_hi, y_lo, y_hi, ks, dt) of the accepted step that crossed t1
    while t1 > t:
        if not (F32(t + dt) > t):
            raise RuntimeError("underflow in dt %r" % dt)
        ks[0] = f
        S_i = S
        y1 = kern.new()
        for i in range(6):
            # stage i produces k_i (already known for i == 0) and the state + support of stage i+1
            row = tab.a[i + 1]
            coefs = [F32(dt * F32(c)) for c in row]
            t_n = F32(t + F32(tab.c[i + 1]) * dt)
            kprev, cprev = _nz(ks[:i], coefs[:i])
            y_next = y1 if i == 5 else Yb[i & 1]
            S_n = Sb[i & 1]
            if i == 0:
                # k_0 = f is known (FSAL): only the combination and the next support are needed
                ops.rk_combine(y, [f], [coefs[0]], out=y_next)
                kern.transform(y_next, t_n, S_n)
            else:
                kern.stage_fwd(S_i, ks[i], y, kprev, cprev, coefs[i], y_next, t_n, S_n)
            S_i = S_n
        kern.stage_fwd(S_i, ks[6])  # k_6 = f(t + dt, y1): FSAL derivative of the next step
        msr = kern.scalar(_msr(kern, y, y1, ks, [F32(dt * F32(c)) for c in tab.c_err], rtol, atol)) / n_el
        accept = msr <= 1.0
        dt_next = _optimal_step(dt, msr)
        if stats is not None:
            key = "accepted" if accept else "rejected"
            stats[key] = stats.get(key, 0) + 1
        if accept:
            t_new = F32(t + dt)
            if t_new >= t1:
                last = (t, t_new, y, y1, list(ks), dt)
            f = ks[6]
            if t_new < t1:
                # S_i (= Sb[1]) already is the support of (y1, t + dt): it becomes the next step's S
                S, Sb[1] = S_i, S
                ks = [ks[6]] + ks[1:6] + [ks[0]]
            t, y = t_new, y1
        dt = dt_next
    t_lo, t_hi, y_lo, y_hi, kk, dts = last
    if t_hi == t1:
        return y_hi
    x = (t1 - t_lo) / (t_hi - t_lo)
    w0, w1, wk = _interp_weights(x, dts, tab.c_mid)
    terms, weights = _nz([y_lo, y_hi] + kk, [w0, w1] + wk)
    return ops.rk_combine(None, terms, weights)


def gcn_solve_adjoint(kern, y1, g1, t0, t1, method="dopri5", step_size=None, rtol=1e-5, atol=1e-5, stats=None):
    """Adjoint pass: integrate (y, a_y, a_t, a_theta) from t1 back to t0.  Returns (a_y(t0), a_theta, a_t)."""
    tab = TABLEAUS[method]
    d = kern.d
    P = kern.n_theta
    dev = kern.dev
    # a_t(t1) = - <f(t1, y1), g>   (torchdiffeq evaluates func once here: counted in nfe)
    S = kern.transform(y1, t1, kern.new_S())
    f1 = kern.new()
    kern.stage_fwd(S, f1)
    # the dot product only feeds a_t, which nothing reads for fixed-step methods (the evaluation itself is
    # kept: the reference's nfe_b counts it)
    a_t = (kern.reduce_small(-(f1 * g1).sum().reshape(1)).reshape(()) if method not in FIXED_METHODS
           else torch.zeros((), dtype=torch.float32, device=dev))
    del f1
    a_theta = torch.zeros(P, dtype=torch.float32, device=dev)
    if method in FIXED_METHODS:
        grid = _grid(t1, t0, step_size)
        y, a = y1, g1
        for g0_, g1_ in zip(grid[:-1], grid[1:]):
            h = F32(g1_ - g0_)
            y, a, S, dth = _gcn_aug_fixed_step(kern, tab, g0_, h, y, a, S, want_S=g1_ != grid[-1],
                                               want_y=g1_ != grid[-1])
            a_theta += dth
            if stats is not None:
                stats["accepted"] = stats.get("accepted", 0) + 1
        kern.reduce_small(a_theta)
        a_t = a_t + a_theta[P - 1]
        return a, a_theta[:P - 1], a_t
    return _gcn_aug_dopri5(kern, tab, y1, g1, a_t, F32(t1), F32(t0), S, rtol, atol, stats)


def _gcn_aug_fixed_step(kern, tab, t0, h, y0, a0, S0, want_S, want_y=True):
    """One explicit RK step of the augmented system with step h (negative: backwards in time).  ``want_y`` False (the
    last step of the grid: nothing reads y(t0) any more) drops the y-combination of the last stage altogether."""
    s = tab.s
    P = kern.n_theta
    ky, ka = [], []
    gth = torch.empty(s, P, dtype=torch.float32, device=kern.dev)
    # two gP buffers in turn: with the peer-memory exchange a push of stage i+1's gP must not land in the buffer a
    # peer's phase 2 of stage i may still be gathering from (peer.py)
    gPs = [kern.new_gP(), kern.new_gP()]
    Ybuf, Abuf = [kern.new_Y(), kern.new_Y()], [kern.new(), kern.new()]
    Y_i, A_i, S_i = y0, a0, S0
    y_out = a_out = S_out = None
    Vy = Va = None
    running = _running_final(tab)
    for i in range(s):
        last = i == s - 1
        t_i = F32(t0 + F32(tab.c[i]) * h)
        t_n = F32(t0 + h) if last else F32(t0 + F32(tab.c[i + 1]) * h)
        base_y0, kyp, cyp, c_self, sec = _stage_plan(tab, i, h, ky, running)
        _, kap, cap, _, _ = _stage_plan(tab, i, h, ka, running)
        sec_y = sec if want_y else None          # without y(t0) the y-state needs no final combination at all
        store_y = (not last) and _needed_later(tab, i) and sec is None
        ky_i = kern.new() if store_y else None
        need_Y = (not last) or want_y
        Y_n = ((kern.new_Y() if want_S else kern.new()) if last else Ybuf[i & 1]) if need_Y else None
        gP = gPs[i & 1]
        second = None
        if sec_y is not None:
            Vy = kern.new()
            second = (sec_y[0], sec_y[1], Vy)
        if need_Y:
            kern.vjp_phase1(S_i, A_i, -1.0, ky_i, gP, y0 if base_y0 else Vy, kyp, cyp, c_self, Y_n, second=second)
        else:
            kern.vjp_phase1(S_i, A_i, -1.0, None, gP)
        # The next stage's support only needs Y_n, which phase 1 has just produced: issue its transform BEFORE
        # phase 2, so that on the row-partitioned path both halo exchanges (gP, then S_n) are in flight underneath
        # the transform and phase 2's dense chain instead of being exposed.
        if (not last) or want_S:
            S_n = kern.transform(Y_n, t_n, kern.new_S())
        else:
            S_n = None
        # k_a of this stage and the adjoint state of the next one in the same pass (k_a is stored only if a later
        # combination reads it; the last-but-one stage stores the running final combination instead)
        ka_i = kern.new() if store_y else None
        A_n = kern.new() if last else Abuf[i & 1]
        second = None
        if sec is not None:
            Va = kern.new()
            second = (sec[0], sec[1], Va)
        kern.vjp_phase2(Y_i, t_i, gP, ka_i, gth[i], a0 if base_y0 else Va, kap, cap, c_self, A_n, second=second)
        ky.append(ky_i)
        ka.append(ka_i)
        Y_i, A_i, S_i = Y_n, A_n, S_n
        y_out, a_out, S_out = Y_n, A_n, S_n
    # weights as host scalars (no host-to-device copy: the step must be capturable in a CUDA graph)
    dth = None
    for i, b in enumerate(tab.b):
        if b != 0:
            term = gth[i] * float(F32(h * F32(b)))
            dth = term if dth is None else dth + term
    return y_out, a_out, S_out, dth


def _gcn_aug_eval(kern, S, y, a, t, ky, ka, gth, gP):
    """F(t, aug) on the fused kernels: (f, -a^T df/dy, [-a^T df/dtheta | -a^T df/dt])."""
    kern.vjp_phase1(S, a, -1.0, ky, gP)
    kern.vjp_phase2(y, t, gP, ka, gth)


def _small_msr(err, v0, v1, rtol, atol):
    tol = atol + rtol * torch.maximum(v0.abs(), v1.abs())
    q = err / tol
    return float((q * q).mean().item())


def _gcn_aug_dopri5(kern, tab, y1, g1, a_t1, t_start, t_end, S, rtol, atol, stats):
    """dopri5 on the augmented system, stepping from t_start (= t1) down to t_end (= t0).

    torchdiffeq integrates the mirrored problem (s = -t, F' = -F); stepping with h < 0 on F is the same
    arithmetic.  The error test is per tensor of the tuple (y, a_y, a_t, a_theta): all four must pass.
    """
    P = kern.n_theta
    dev = kern.dev
    n_el = kern.numel_global()
    new = kern.new
    theta0 = torch.zeros(P - 1, dtype=torch.float32, device=dev)   # a_theta(t1)
    at0 = a_t1.reshape(1).to(torch.float32)
    gP = kern.new_gP()

    def rms_big(x, ref):
        return float(np.sqrt(kern.scalar(ops.rk_error_sumsq(ref, ref, [x], [1.0], rtol, atol)) / n_el))

    def rms_small(x, ref):
        return float(torch.sqrt(((x / (atol + rtol * ref.abs())) ** 2).mean()).item())

    # f0 = F(t1, aug0)
    ky0, ka0 = new(), new()
    g0 = torch.empty(P, dtype=torch.float32, device=dev)
    _gcn_aug_eval(kern, S, y1, g1, t_start, ky0, ka0, g0, gP)
    kern.reduce_small(g0)
    d0s = [rms_big(y1, y1), rms_big(g1, g1), rms_small(at0, at0), rms_small(theta0, theta0)]
    d1s = [rms_big(ky0, y1), rms_big(ka0, g1), rms_small(g0[P - 1:], at0), rms_small(g0[:P - 1], theta0)]
    d1 = max(d1s)
    # torchdiffeq works on the mirrored problem: its f0 is -F and its step is positive; norms are identical
    h0 = _initial_h0(d0s, d1s)
    yp, ap, Sp, kyp, kap = new(), new(), kern.new_S(), new(), new()
    gp_ = torch.empty(P, dtype=torch.float32, device=dev)
    ops.rk_combine(y1, [ky0], [-h0], out=yp)
    ops.rk_combine(g1, [ka0], [-h0], out=ap)
    kern.transform(yp, F32(t_start - h0), Sp)
    # (a_t, a_theta do not enter F, so the probe state only needs y and a_y)
    _gcn_aug_eval(kern, Sp, yp, ap, F32(t_start - h0), kyp, kap, gp_, gP)
    kern.reduce_small(gp_)
    d2 = max(
        float(np.sqrt(kern.scalar(ops.rk_error_sumsq(y1, y1, [kyp, ky0], [1.0, -1.0], rtol, atol)) / n_el)),
        float(np.sqrt(kern.scalar(ops.rk_error_sumsq(g1, g1, [kap, ka0], [1.0, -1.0], rtol, atol)) / n_el)),
        rms_small(gp_[P - 1:] - g0[P - 1:], at0), rms_small(gp_[:P - 1] - g0[:P - 1], theta0)) / float(h0)
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(F32(1e-6), F32(h0 * F32(1e-3)))
    else:
        h1 = F32((0.01 / max(d1, d2)) ** (1.0 / 5.0))
    dt = F32(min(F32(100.0) * h0, h1))   # magnitude; the step taken is h = -dt
    del yp, ap, Sp, kyp, kap

    t = t_start
    y, a, at, th = y1, g1, at0, theta0
    ky = [ky0] + [new() for _ in range(6)]
    ka = [ka0] + [new() for _ in range(6)]
    gth = torch.empty(7, P, dtype=torch.float32, device=dev)
    gth[0] = g0
    Yb, Ab, Ss = [new(), new()], [new(), new()], kern.new_S()
    last = None
    while t > t_end:
        if not (F32(t - dt) < t):
            raise RuntimeError("underflow in dt %r" % dt)
        h = F32(-dt)
        S_i, Y_i, A_i = S, y, a
        y_new, a_new = new(), new()
        for i in range(6):
            row = tab.a[i + 1]
            coefs = [F32(h * F32(c)) for c in row]
            t_i = F32(t + F32(tab.c[i]) * h)
            t_n = F32(t + F32(tab.c[i + 1]) * h)
            Y_n = y_new if i == 5 else Yb[i & 1]
            A_n = a_new if i == 5 else Ab[i & 1]
            if i == 0:
                ops.rk_combine(y, [ky[0]], [coefs[0]], out=Y_n)
                kern.transform(Y_n, t_n, Ss)
            else:
                kyp_, cyp_ = _nz(ky[:i], coefs[:i])
                kern.vjp_phase1(S_i, A_i, -1.0, ky[i], gP, y, kyp_, cyp_, coefs[i], Y_n)
                kern.transform(Y_n, t_n, Ss)      # before phase 2: see _gcn_aug_fixed_step (S_i is Ss or S, both consumed)
                kern.vjp_phase2(Y_i, t_i, gP, ka[i], gth[i])
            kap_, cap_ = _nz(ka[:i + 1], coefs)
            ops.rk_combine(a, kap_, cap_, out=A_n)
            S_i, Y_i, A_i = Ss, Y_n, A_n
        _gcn_aug_eval(kern, Ss, y_new, a_new, F32(t + h), ky[6], ka[6], gth[6], gP)
        kern.reduce_small(gth[1:])     # row 0 is carried over from the previous step (FSAL), already summed
        cerr = [F32(h * F32(c)) for c in tab.c_err]
        wb = torch.tensor([float(F32(h * F32(b))) for b in tab.b], dtype=torch.float32, device=dev)
        we = torch.tensor([float(c) for c in cerr], dtype=torch.float32, device=dev)
        small_new = (wb[:, None] * gth).sum(0)
        small_err = (we[:, None] * gth).sum(0)
        th_new, at_new = th + small_new[:P - 1], at + small_new[P - 1:]
        msrs = (kern.scalar(_msr(kern, y, y_new, ky, cerr, rtol, atol)) / n_el,
                kern.scalar(_msr(kern, a, a_new, ka, cerr, rtol, atol)) / n_el,
                _small_msr(small_err[P - 1:], at, at_new, rtol, atol),
                _small_msr(small_err[:P - 1], th, th_new, rtol, atol))
        msr = max(msrs)
        accept = msr <= 1.0
        dt_next = _optimal_step(dt, msr)
        if stats is not None:
            key = "accepted" if accept else "rejected"
            stats[key] = stats.get(key, 0) + 1
            if "trace" in stats:     # per-step diagnostics (tests/d64_noise.py): (t, dt, per-tensor error ratios)
                stats["trace"].append((float(t), float(dt), [float(m) for m in msrs]))
        if accept:
            t_new = F32(t + h)
            if t_new <= t_end:
                last = (t, t_new, (y, a, at, th), (y_new, a_new, at_new, th_new), list(ky), list(ka), gth.clone(), h)
            t, y, a, at, th = t_new, y_new, a_new, at_new, th_new
            if t > t_end:
                S, Ss = Ss, S   # Ss already holds the support of (y_new, t + h)
                ky = [ky[6]] + ky[1:6] + [ky[0]]
                ka = [ka[6]] + ka[1:6] + [ka[0]]
                gth[0] = gth[6]
        dt = dt_next
    t_lo, t_hi, lo, hi, kyl, kal, gl, hs = last
    if t_hi == t_end:
        return hi[1], hi[3], hi[2]
    x = (t_end - t_lo) / (t_hi - t_lo)
    w0, w1, wk = _interp_weights(x, hs, tab.c_mid)
    ta, wa = _nz([lo[1], hi[1]] + kal, [w0, w1] + wk)
    a_out = ops.rk_combine(None, ta, wa)
    wkt = torch.tensor(wk, dtype=torch.float32, device=dev)
    small = (wkt[:, None] * gl).sum(0)
    th_out = w0 * lo[3] + w1 * hi[3] + small[:P - 1]
    at_out = w0 * lo[2] + w1 * hi[2] + small[P - 1:]
    return a_out, th_out, at_out


def _make_kernel(plan, *args):
    """GcnKernel for a single-device plan; the plan's own kernel class for a row-partitioned one."""
    factory = getattr(plan, "make_kernel", None)
    return factory(*args) if factory is not None else GcnKernel(plan, *args)


class _GcnAdjointFn(torch.autograd.Function):
    """autograd node for the fused GCN ODE block: forward solve without a tape, adjoint solve in backward."""

    @staticmethod
    def forward(ctx, y0, weight, bias, gamma, beta, funcmod, plan, t0, t1, rtol, atol, method, step_size, stats):
        y0 = ops._rowmajor(y0, "y0")
        kern = _make_kernel(plan, weight, bias, gamma, beta, funcmod.norm1.num_groups, funcmod.norm1.eps,
                            getattr(funcmod, "precision", _lib.PREC_FP32))
        y1 = gcn_solve_forward(kern, y0, t0, t1, method, step_size, rtol, atol,
                               None if stats is None else stats.setdefault("forward", {}))
        funcmod.nfe += kern.nfe
        ctx.funcmod, ctx.plan, ctx.cfg, ctx.stats = funcmod, plan, (t0, t1, rtol, atol, method, step_size), stats
        ctx.has_bias = bias is not None
        ctx.save_for_backward(y1, weight, bias, gamma, beta)
        return y1

    @staticmethod
    def backward(ctx, g):
        y1, weight, bias, gamma, beta = ctx.saved_tensors
        t0, t1, rtol, atol, method, step_size = ctx.cfg
        fm = ctx.funcmod
        kern = _make_kernel(ctx.plan, weight, bias, gamma, beta, fm.norm1.num_groups, fm.norm1.eps,
                            getattr(fm, "precision", _lib.PREC_FP32))
        g = ops._rowmajor(g, "grad").contiguous()
        a, th, _ = gcn_solve_adjoint(kern, y1, g, t0, t1, method, step_size, rtol, atol,
                                     None if ctx.stats is None else ctx.stats.setdefault("backward", {}))
        fm.nfe += kern.nfe
        d = kern.d
        nw = (d + 1) * d
        gw = th[:nw].reshape(d + 1, d)
        gb = th[nw:nw + d] if ctx.has_bias else None
        gg = th[nw + d:nw + 2 * d]
        gbeta = th[nw + 2 * d:nw + 3 * d]
        return (a, gw, gb, gg, gbeta) + (None,) * 9


# ---------------------------------------------------------------------------------------------
# generic engine (tuple states, module forward + autograd for the adjoint)
# ---------------------------------------------------------------------------------------------


def _is_big(x):
    return x.is_cuda and x.dtype == torch.float32 and x.numel() >= 4096 and x.is_contiguous()


def _lincomb(y0, ks, coefs):
    ks, coefs = _nz(ks, coefs)
    if not ks:
        return y0.clone()
    if _is_big(ks[0]) and all(k.is_contiguous() for k in ks) and (y0 is None or y0.is_contiguous()):
        return ops.rk_combine(y0, ks, coefs)
    tot = None
    for k, c in zip(ks, coefs):
        tot = float(c) * k if tot is None else tot + float(c) * k
    return tot if y0 is None else y0 + tot


def _ratio_msr(y0, y1, ks, coefs, rtol, atol):
    ks_, cs_ = _nz(ks, coefs)
    if _is_big(y0) and y1.is_contiguous() and all(k.is_contiguous() for k in ks_):
        return float(ops.rk_error_sumsq(y0, y1, ks_, cs_, rtol, atol).item()) / y0.numel()
    err = _lincomb(None, ks_, cs_) if ks_ else torch.zeros_like(y0)
    return _small_msr(err, y0, y1, rtol, atol)


def _rms_scaled(x, ref, rtol, atol):
    if _is_big(ref) and x.is_contiguous():
        return float(np.sqrt(ops.rk_error_sumsq(ref, ref, [x], [1.0], rtol, atol).item() / ref.numel()))
    return float(torch.sqrt(((x / (atol + rtol * ref.abs())) ** 2).mean()).item())


def _generic_solve(func, y0, t0, t1, method, step_size, rtol, atol, stats):
    """y0: tuple of tensors; func(t_tensor, tuple) -> tuple.  Returns the tuple at t1."""
    tab = TABLEAUS[method]
    n = len(y0)
    dev = y0[0].device

    def ft(tt, yy):
        return tuple(func(torch.tensor(float(tt), dtype=y0[0].dtype, device=dev), yy))

    def rk_stages(t, h, y, f0=None):
        ks = [[] for _ in range(n)]
        for i in range(tab.s):
            if i == 0:
                fi = f0 if f0 is not None else ft(t, y)
            else:
                coefs = [F32(h * F32(c)) for c in tab.a[i]]
                yi = tuple(_lincomb(y[q], ks[q], coefs) for q in range(n))
                fi = ft(F32(t + F32(tab.c[i]) * h), yi)
            for q in range(n):
                ks[q].append(fi[q])
        return ks

    if method in FIXED_METHODS:
        grid = _grid(t0, t1, step_size)
        y = tuple(y0)
        for g0, g1 in zip(grid[:-1], grid[1:]):
            h = F32(g1 - g0)
            ks = rk_stages(g0, h, y)
            wb = [F32(h * F32(b)) for b in tab.b]
            y = tuple(_lincomb(y[q], ks[q], wb) for q in range(n))
            if stats is not None:
                stats["accepted"] = stats.get("accepted", 0) + 1
        return y

    # dopri5 (direction-aware: sgn = -1 integrates backwards)
    sgn = F32(1.0) if t1 >= t0 else F32(-1.0)
    t = F32(t0)
    y = tuple(y0)
    f = ft(t, y)
    d0s = [_rms_scaled(y[q], y[q], rtol, atol) for q in range(n)]
    d1s = [_rms_scaled(f[q], y[q], rtol, atol) for q in range(n)]
    d1 = max(d1s)
    h0 = _initial_h0(d0s, d1s)
    yp = tuple(_lincomb(y[q], [f[q]], [sgn * h0]) for q in range(n))
    fp = ft(F32(t + sgn * h0), yp)
    d2 = max(_rms_scaled(fp[q] - f[q], y[q], rtol, atol) for q in range(n)) / float(h0)
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(F32(1e-6), F32(h0 * F32(1e-3)))
    else:
        h1 = F32((0.01 / max(d1, d2)) ** (1.0 / 5.0))
    dt = F32(min(F32(100.0) * h0, h1))
    last = None
    while (t1 - t) * sgn > 0:
        h = F32(sgn * dt)
        if not (F32(t + h) != t):
            raise RuntimeError("underflow in dt %r" % dt)
        ks = rk_stages(t, h, y, f0=f)
        wb = [F32(h * F32(b)) for b in tab.b]
        y_new = tuple(_lincomb(y[q], ks[q][:6], wb[:6]) for q in range(n))
        # stage 6 of the FSAL tableau was evaluated at exactly y_new
        cerr = [F32(h * F32(c)) for c in tab.c_err]
        msr = max(_ratio_msr(y[q], y_new[q], ks[q], cerr, rtol, atol) for q in range(n))
        accept = msr <= 1.0
        if stats is not None:
            key = "accepted" if accept else "rejected"
            stats[key] = stats.get(key, 0) + 1
        if accept:
            t_new = F32(t + h)
            if (t1 - t_new) * sgn <= 0:
                last = (t, t_new, y, y_new, ks, h)
            t, y, f = t_new, y_new, tuple(ks[q][6] for q in range(n))
        dt = _optimal_step(dt, msr)
    t_lo, t_hi, y_lo, y_hi, ks, hs = last
    if t_hi == F32(t1):
        return y_hi
    x = (F32(t1) - t_lo) / (t_hi - t_lo)
    w0, w1, wk = _interp_weights(x, hs, tab.c_mid)
    return tuple(_lincomb(None, [y_lo[q], y_hi[q]] + ks[q], [w0, w1] + wk) for q in range(n))


class _GenericAdjointFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, func, cfg, stats, n_y, *ys_and_params):
        t0, t1, rtol, atol, method, step_size = cfg
        y0 = tuple(ys_and_params[:n_y])
        with torch.no_grad():
            y1 = _generic_solve(func, y0, t0, t1, method, step_size, rtol, atol,
                                None if stats is None else stats.setdefault("forward", {}))
        ctx.func, ctx.cfg, ctx.stats, ctx.n_y = func, cfg, stats, n_y
        ctx.save_for_backward(*y1)
        return tuple(y1)

    @staticmethod
    def backward(ctx, *grads):
        func, n = ctx.func, ctx.n_y
        t0, t1, rtol, atol, method, step_size = ctx.cfg
        y1 = ctx.saved_tensors
        params = tuple(p for p in func.parameters() if p.requires_grad)
        dev = y1[0].device

        def aug(tt, state):
            y, a = state[:n], state[n:2 * n]
            with torch.enable_grad():
                tt_ = tt.detach().requires_grad_(True)
                y_ = tuple(v.detach().requires_grad_(True) for v in y)
                f = func(tt_, y_)
                f = (f,) if torch.is_tensor(f) else tuple(f)
                vj = torch.autograd.grad(f, (tt_,) + y_ + params, tuple(-v for v in a), allow_unused=True)
            vt = torch.zeros(1, device=dev, dtype=y[0].dtype) if vj[0] is None else vj[0].reshape(1)
            vy = tuple(torch.zeros_like(v) if g is None else g.contiguous() for g, v in zip(vj[1:1 + n], y_))
            flat = [torch.zeros_like(p).reshape(-1) if g is None else g.reshape(-1) for g, p in zip(vj[1 + n:], params)]
            vp = torch.cat(flat) if flat else torch.zeros(1, device=dev, dtype=y[0].dtype)
            return tuple(v.detach() for v in f) + vy + (vt, vp)

        with torch.no_grad():
            a = tuple(g.contiguous() if g is not None else torch.zeros_like(v) for g, v in zip(grads, y1))
            tt = torch.tensor(float(t1), dtype=y1[0].dtype, device=dev)
            f1 = func(tt, tuple(y1))
            f1 = (f1,) if torch.is_tensor(f1) else tuple(f1)
            a_t = -sum((u * v).sum() for u, v in zip(f1, a)).reshape(1)
            a_p = torch.zeros(max(sum(p.numel() for p in params), 1), device=dev, dtype=y1[0].dtype)
            out = _generic_solve(aug, tuple(y1) + a + (a_t, a_p), t1, t0, method, step_size, rtol, atol,
                                 None if ctx.stats is None else ctx.stats.setdefault("backward", {}))
        a0 = out[n:2 * n]
        a_p = out[2 * n + 1]
        gp, off = [], 0
        for p in func.parameters():
            if p.requires_grad:
                gp.append(a_p[off:off + p.numel()].reshape(p.shape))
                off += p.numel()
            else:
                gp.append(None)
        return (None, None, None, None) + tuple(a0) + tuple(gp)


# ---------------------------------------------------------------------------------------------
# public API
# ---------------------------------------------------------------------------------------------


def _cfg(t, rtol, atol, method, options):
    ts = _times(t)
    if len(ts) != 2:
        raise NotImplementedError("graph-odenet integrates over t = [t0, t1] (GCN/models.py:186); got %d points" % len(ts))
    method = method or "dopri5"
    if method not in TABLEAUS:
        raise ValueError("unsupported method %r (have %s)" % (method, sorted(TABLEAUS)))
    step = (options or {}).get("step_size")
    return ts[0], ts[1], float(rtol), float(atol), method, step


def odeint_adjoint(func, y0, t, rtol=1e-6, atol=1e-12, method=None, options=None, stats=None):
    """torchdiffeq.odeint_adjoint for t = [t0, t1]: returns a tensor stacked over t (out[1] = y(t1))."""
    t0, t1, rtol, atol, method, step = _cfg(t, rtol, atol, method, options)
    fused = getattr(func, "_gode_fused", None)
    plan = fused() if (fused is not None and torch.is_tensor(y0)) else None
    if plan is not None:
        gc = func.gc1
        y1 = _GcnAdjointFn.apply(y0, gc.weight, gc.bias, func.norm1.weight, func.norm1.bias, func, plan,
                                 t0, t1, rtol, atol, method, step, stats)
        return torch.stack([y0, y1])
    tensor_in = torch.is_tensor(y0)
    ys = (y0,) if tensor_in else tuple(y0)
    params = tuple(func.parameters())
    wrapped = func if not tensor_in else _TensorFunc(func)
    out = _GenericAdjointFn.apply(wrapped, (t0, t1, rtol, atol, method, step), stats, len(ys), *ys, *params)
    stacked = tuple(torch.stack([a, b]) for a, b in zip(ys, out))
    return stacked[0] if tensor_in else stacked


def odeint_adjoint_final(func, y0, t, rtol=1e-6, atol=1e-12, method=None, options=None, stats=None):
    """Same solve as ``odeint_adjoint`` but returns only y(t1) (no [2, N, d] stack is materialised)."""
    t0, t1, rtol, atol, method, step = _cfg(t, rtol, atol, method, options)
    fused = getattr(func, "_gode_fused", None)
    plan = fused() if fused is not None else None
    if plan is not None:
        gc = func.gc1
        return _GcnAdjointFn.apply(y0, gc.weight, gc.bias, func.norm1.weight, func.norm1.bias, func, plan,
                                   t0, t1, rtol, atol, method, step, stats)
    params = tuple(func.parameters())
    out = _GenericAdjointFn.apply(_TensorFunc(func), (t0, t1, rtol, atol, method, step), stats, 1, y0, *params)
    return out[0]


class _TensorFunc(torch.nn.Module):
    """Presents a tensor-state ODE function to the tuple-state engine."""

    def __init__(self, base):
        super().__init__()
        self.base = base

    def forward(self, t, y):
        if isinstance(y, tuple):
            return (self.base(t, y[0]),)
        return self.base(t, y)


def odeint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None, stats=None):
    """torchdiffeq.odeint (no adjoint: differentiates through the solver's own operations)."""
    t0, t1, rtol, atol, method, step = _cfg(t, rtol, atol, method, options)
    tensor_in = torch.is_tensor(y0)
    ys = (y0,) if tensor_in else tuple(y0)
    f = (lambda tt, yy: (func(tt, yy[0]),)) if tensor_in else func
    out = _generic_solve(f, ys, t0, t1, method, step, rtol, atol, stats)
    stacked = tuple(torch.stack([a, b]) for a, b in zip(ys, out))
    return stacked[0] if tensor_in else stacked
