"""Restatement of ``torchdiffeq.odeint`` / ``odeint_adjoint`` (TEST INFRASTRUCTURE, CPU oracle).

The reference calls ``from torchdiffeq import odeint_adjoint as odeint`` (GCN/models.py:5) and
``odeint(self.odefunc, x, self.integration_time, rtol=self.tol, atol=self.tol)`` (GCN/models.py:192,
GAT/models.py:192).  torchdiffeq (github.com/rtqichen/torchdiffeq) is third-party, un-vendored,
un-pinned and not installable in this image => **PARITY UNPINNED** for this file.  What is restated
here is the published algorithm of the 0.0.x line (contemporary with the 2019 reference):

* fixed-grid ``euler`` / ``midpoint`` / ``rk4`` (rk4 is the 3/8-rule "alt" step), grid = ``t`` itself or
  ``options={'step_size': h}``, outputs by linear interpolation inside a step;
* adaptive ``dopri5`` (Dormand-Prince 5(4), FSAL, Hairer initial step with RMS norm, per-tensor
  mean-square error ratio, safety 0.9 / ifactor 10 / dfactor 0.2, quartic dense output, the solver
  steps past the requested time and interpolates);
* the adjoint: forward under ``no_grad``; backward integrates the augmented state
  ``(y, a_y, a_t, a_theta)`` from t1 to t0 with the same method and tolerances (decreasing time is
  handled by negating ``t`` and ``f``).

Pinned nevertheless, against an independent implementation in the image (tests/test_oracle_golden.py::
test_solver_pieces_against_scipy, float64, 1e-14): one Dormand-Prince step (stages, solution, FSAL derivative) vs scipy's RK45
``rk_step``; the initial step vs scipy's ``select_initial_step``; the fixed-step increments vs the textbook formulas.  NOT
pinned: torchdiffeq's own error coefficients (``_DP_CERR`` ends in -1/60, Dormand-Prince's pair in -1/40), controller and
dense output.

States are tuples of tensors throughout, exactly as in the package.  ``stats`` (optional dict) receives
accepted/rejected step counts so adaptive solves can be compared "on the accepted-step count".
"""
from __future__ import annotations

import torch

# ----------------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------------


def _rms(x):
    return x.norm() / (x.numel() ** 0.5)


def _tuple(x):
    return (x,) if torch.is_tensor(x) else tuple(x)


def _scaled_dot(scale, coeffs, ks):
    """sum_j (scale*coeffs[j]) * ks[j], evaluated left to right (python ``sum`` order)."""
    tot = 0
    for c, k in zip(coeffs, ks):
        tot = tot + (scale * c) * k
    return tot


def _dot(coeffs, ks):
    tot = 0
    for c, k in zip(coeffs, ks):
        tot = tot + c * k
    return tot


# ----------------------------------------------------------------------------------------------
# fixed grid
# ----------------------------------------------------------------------------------------------


def _euler_step(func, t, dt, y):
    return tuple(dt * f_ for f_ in func(t, y))


def _midpoint_step(func, t, dt, y):
    y_mid = tuple(y_ + f_ * dt / 2 for y_, f_ in zip(y, func(t, y)))
    return tuple(dt * f_ for f_ in func(t + dt / 2, y_mid))


def _rk4_38_step(func, t, dt, y):
    """3/8-rule fourth order step (what torchdiffeq's ``method='rk4'`` uses)."""
    k1 = func(t, y)
    k2 = func(t + dt / 3, tuple(y_ + dt * k1_ / 3 for y_, k1_ in zip(y, k1)))
    k3 = func(t + dt * 2 / 3, tuple(y_ + dt * (k2_ - k1_ / 3) for y_, k1_, k2_ in zip(y, k1, k2)))
    k4 = func(t + dt, tuple(y_ + dt * (k1_ - k2_ + k3_) for y_, k1_, k2_, k3_ in zip(y, k1, k2, k3)))
    return tuple((k1_ + 3 * (k2_ + k3_) + k4_) * dt / 8 for k1_, k2_, k3_, k4_ in zip(k1, k2, k3, k4))


_FIXED = {"euler": _euler_step, "midpoint": _midpoint_step, "rk4": _rk4_38_step}


def _fixed_grid(t, step_size):
    if step_size is None:
        return t
    start, end = t[0], t[-1]
    niters = int(torch.ceil((end - start) / step_size + 1).item())
    grid = torch.arange(0, niters, dtype=t.dtype) * step_size + start
    if grid[-1] > t[-1]:
        grid[-1] = t[-1]
    return grid


def _solve_fixed(func, y0, t, method, step_size, stats):
    step = _FIXED[method]
    grid = _fixed_grid(t, step_size)
    out = [y0]
    j = 1
    y = y0
    for t0, t1 in zip(grid[:-1], grid[1:]):
        dy = step(func, t0, t1 - t0, y)
        y1 = tuple(a + b for a, b in zip(y, dy))
        while j < len(t) and t1 >= t[j]:
            if t[j] == t0:
                out.append(y)
            elif t[j] == t1:
                out.append(y1)
            else:
                out.append(tuple(a + (b - a) / (t1 - t0) * (t[j] - t0) for a, b in zip(y, y1)))
            j += 1
        y = y1
        if stats is not None:
            stats["accepted"] = stats.get("accepted", 0) + 1
    return out


# ----------------------------------------------------------------------------------------------
# dopri5
# ----------------------------------------------------------------------------------------------

_DP_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_DP_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_DP_CSOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
_DP_CERR = [
    35 / 384 - 1951 / 21600,
    0,
    500 / 1113 - 22642 / 50085,
    125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400,
    11 / 84 - 649 / 6300,
    -1.0 / 60.0,
]
_DP_CMID = [
    6025192743 / 30085553152 / 2,
    0,
    51252292925 / 65400821598 / 2,
    -2691868925 / 45128329728 / 2,
    187940372067 / 1594534317056 / 2,
    -1776094331 / 19743644256 / 2,
    11237099 / 235043384 / 2,
]


def _initial_step(func, t0, y0, order, rtol, atol, f0):
    scale = tuple(a + torch.abs(y_) * r for y_, a, r in zip(y0, atol, rtol))
    d0 = tuple(_rms(y_ / s_) for y_, s_ in zip(y0, scale))
    d1 = tuple(_rms(f_ / s_) for f_, s_ in zip(f0, scale))
    if max(d0).item() < 1e-5 or max(d1).item() < 1e-5:
        h0 = torch.tensor(1e-6).to(t0)
    else:
        h0 = 0.01 * max(a / b for a, b in zip(d0, d1))
    y1 = tuple(y_ + h0 * f_ for y_, f_ in zip(y0, f0))
    f1 = func(t0 + h0, y1)
    d2 = tuple(_rms((a - b) / s_) / h0 for a, b, s_ in zip(f1, f0, scale))
    if max(d1).item() <= 1e-15 and max(d2).item() <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6).to(h0), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1 + d2)) ** (1.0 / float(order + 1))
    return torch.min(100 * h0, h1)


def _optimal_step(last, msr, safety=0.9, ifactor=10.0, dfactor=0.2, order=5):
    msr = max(msr)
    if msr == 0:
        return last * ifactor
    if msr < 1:
        dfactor = 1.0
    ratio = torch.sqrt(msr).to(last)
    factor = torch.max(
        torch.tensor(1 / ifactor).to(last),
        torch.min(ratio ** (1.0 / order) / safety, torch.tensor(1 / dfactor).to(last)),
    )
    return last / factor


def _dp_step(func, y0, f0, t0, dt):
    k = [[f_] for f_ in f0]  # per tensor list of stages
    for alpha_i, beta_i in zip(_DP_ALPHA, _DP_BETA):
        ti = t0 + alpha_i * dt
        yi = tuple(y_ + _scaled_dot(dt, beta_i, k_) for y_, k_ in zip(y0, k))
        fi = func(ti, yi)
        for k_, f_ in zip(k, fi):
            k_.append(f_)
    # FSAL tableau: c_sol[:-1] == beta[-1] and c_sol[-1] == 0, so yi already is y1
    y1 = yi
    f1 = tuple(k_[-1] for k_ in k)
    err = tuple(_scaled_dot(dt, _DP_CERR, k_) for k_ in k)
    return y1, f1, err, k


def _interp_fit(y0, y1, k, dt):
    coeffs = []
    for y0_, y1_, k_ in zip(y0, y1, k):
        ym = y0_ + _scaled_dot(dt, _DP_CMID, k_)
        f0_, f1_ = k_[0], k_[-1]
        a = _dot([2 * dt, -2 * dt, -8, -8, 16], [f1_, f0_, y1_, y0_, ym])
        b = _dot([5 * dt, -3 * dt, 18, 14, -32], [f0_, f1_, y0_, y1_, ym])
        c = _dot([dt, -4 * dt, -11, -5, 16], [f1_, f0_, y0_, y1_, ym])
        d = dt * f0_
        e = y0_
        coeffs.append([a, b, c, d, e])
    return coeffs


def _interp_eval(coeffs, t0, t1, t):
    x = ((t - t0) / (t1 - t0)).to(coeffs[0][0].dtype)
    xs = [torch.tensor(1).to(x), x]
    for _ in range(2, 5):
        xs.append(xs[-1] * x)
    return tuple(_dot(c_, list(reversed(xs))) for c_ in coeffs)


def _solve_dopri5(func, y0, t, rtol, atol, stats, max_steps=2 ** 31 - 1):
    n = len(y0)
    rtol = tuple([rtol] * n) if not isinstance(rtol, (tuple, list)) else tuple(rtol)
    atol = tuple([atol] * n) if not isinstance(atol, (tuple, list)) else tuple(atol)
    f0 = func(t[0].type_as(y0[0]), y0)
    dt = _initial_step(func, t[0], y0, 4, rtol, atol, f0).to(t)
    interp = [[y_] * 5 for y_ in y0]
    state = [y0, f0, t[0], t[0], dt, interp]  # y1, f1, t0, t1, dt, interp
    out = [y0]
    for i in range(1, len(t)):
        nsteps = 0
        while t[i] > state[3]:
            assert nsteps < max_steps
            y_, f_, _, t0_, dt_, interp_ = state
            assert t0_ + dt_ > t0_, "underflow in dt"
            y1, f1, err, k = _dp_step(func, y_, f_, t0_, dt_)
            msr = []
            for e_, a_, r_, ya_, yb_ in zip(err, atol, rtol, y_, y1):
                tol = a_ + r_ * torch.max(torch.abs(ya_), torch.abs(yb_))
                q = e_ / tol
                msr.append(torch.mean(q * q))
            accept = all(bool(m <= 1) for m in msr)
            if stats is not None:
                key = "accepted" if accept else "rejected"
                stats[key] = stats.get(key, 0) + 1
                if "trace" in stats:     # per-step diagnostics (tests/d64_noise.py): (t, dt, per-tensor error ratios)
                    stats["trace"].append((float(t0_), float(dt_), [float(m) for m in msr]))
            dt_next = _optimal_step(dt_, msr)
            if accept:
                state = [y1, f1, t0_, t0_ + dt_, dt_next, _interp_fit(y_, y1, k, dt_)]
            else:
                state = [y_, f_, t0_, t0_, dt_next, interp_]
            nsteps += 1
        out.append(_interp_eval(state[5], state[2], state[3], t[i]))
    return out


# ----------------------------------------------------------------------------------------------
# public API
# ----------------------------------------------------------------------------------------------


def odeint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None, stats=None):
    """``torchdiffeq.odeint`` restated.  Returns the solution stacked over ``t`` (tuple in -> tuple out)."""
    tensor_in = torch.is_tensor(y0)
    y0 = _tuple(y0)
    base = func
    if tensor_in:
        func = lambda tt, yy: (base(tt, yy[0]),)  # noqa: E731
    if bool((t[1:] < t[:-1]).all()) and len(t) > 1:  # decreasing -> integrate the mirrored problem
        inner = func
        func = lambda tt, yy: tuple(-f_ for f_ in inner(-tt, yy))  # noqa: E731
        t = -t
    method = method or "dopri5"
    if method == "dopri5":
        sol = _solve_dopri5(func, y0, t, rtol, atol, stats)
    elif method in _FIXED:
        sol = _solve_fixed(func, y0, t, method, (options or {}).get("step_size"), stats)
    else:
        raise ValueError("unsupported method %r" % (method,))
    stacked = tuple(torch.stack([s[i] for s in sol]) for i in range(len(y0)))
    return stacked[0] if tensor_in else stacked


class _Adjoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, func, t, rtol, atol, method, options, stats, n_y, *ys_and_params):
        y0 = tuple(ys_and_params[:n_y])
        ctx.func, ctx.rtol, ctx.atol, ctx.method, ctx.options, ctx.stats = func, rtol, atol, method, options, stats
        ctx.n_y = n_y
        with torch.no_grad():
            ans = odeint(func, y0, t, rtol=rtol, atol=atol, method=method, options=options,
                         stats=None if stats is None else stats.setdefault("forward", {}))
        ctx.save_for_backward(t, *ans)
        return ans

    @staticmethod
    def backward(ctx, *grad_out):
        t, *ans = ctx.saved_tensors
        func, n = ctx.func, ctx.n_y
        params = tuple(p for p in func.parameters()) if hasattr(func, "parameters") else ()

        def aug_dynamics(tt, aug):
            y, a_y = aug[:n], aug[n:2 * n]
            with torch.enable_grad():
                tt_ = tt.to(y[0].device).detach().requires_grad_(True)
                y_ = tuple(v.detach().requires_grad_(True) for v in y)
                f = func(tt_, y_)
                vjps = torch.autograd.grad(f, (tt_,) + y_ + params, tuple(-a for a in a_y),
                                           allow_unused=True, retain_graph=True)
            vjp_t = torch.zeros_like(tt_) if vjps[0] is None else vjps[0]
            vjp_y = tuple(torch.zeros_like(v) if g is None else g for g, v in zip(vjps[1:1 + n], y_))
            flat = [torch.zeros_like(p).reshape(-1) if g is None else g.reshape(-1)
                    for g, p in zip(vjps[1 + n:], params)]
            vjp_p = torch.cat(flat) if flat else torch.tensor(0.0).to(vjp_y[0])
            return (*f, *vjp_y, vjp_t, vjp_p)

        with torch.no_grad():
            a_y = tuple(g[-1] for g in grad_out)
            a_p = torch.zeros(sum(p.numel() for p in params)).to(a_y[0]) if params else torch.tensor(0.0).to(a_y[0])
            a_t = torch.tensor(0.0).to(t)
            time_vjps = []
            for i in range(len(t) - 1, 0, -1):
                y_i = tuple(a[i] for a in ans)
                g_i = tuple(g[i] for g in grad_out)
                f_i = func(t[i], y_i)
                dLdt = sum(torch.dot(a.reshape(-1), b.reshape(-1)).reshape(1) for a, b in zip(f_i, g_i))
                a_t = a_t - dLdt
                time_vjps.append(dLdt)
                aug0 = (*y_i, *a_y, a_t, a_p)
                aug = odeint(aug_dynamics, aug0, torch.stack([t[i], t[i - 1]]), rtol=ctx.rtol, atol=ctx.atol,
                             method=ctx.method, options=ctx.options,
                             stats=None if ctx.stats is None else ctx.stats.setdefault("backward", {}))
                a_y = tuple(a[1] for a in aug[n:2 * n])
                a_t = aug[2 * n][1]
                a_p = aug[2 * n + 1][1]
                a_y = tuple(a + g[i - 1] for a, g in zip(a_y, grad_out))
            time_vjps.append(a_t)
        # gradient wrt params is returned unflattened, in func.parameters() order
        gp, off = [], 0
        for p in params:
            gp.append(a_p[off:off + p.numel()].reshape(p.shape))
            off += p.numel()
        return (None, None, None, None, None, None, None, None, *a_y, *gp)


def odeint_adjoint(func, y0, t, rtol=1e-6, atol=1e-12, method=None, options=None, stats=None):
    """``torchdiffeq.odeint_adjoint`` restated (``func`` must be an ``nn.Module`` or have ``parameters()``)."""
    tensor_in = torch.is_tensor(y0)
    y0t = _tuple(y0)
    if tensor_in:
        class _Wrap(torch.nn.Module):
            def __init__(self, base):
                super().__init__()
                self.base = base

            def parameters(self, recurse=True):
                return self.base.parameters()

            def forward(self, tt, yy):
                return (self.base(tt, yy[0]),)

        f = _Wrap(func)
    else:
        f = func
    params = tuple(f.parameters())
    out = _Adjoint.apply(f, t, rtol, atol, method, options, stats, len(y0t), *y0t, *params)
    return out[0] if tensor_in else out
