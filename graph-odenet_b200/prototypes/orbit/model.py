"""The Interaction-Network models of ``prototypes/orbit/model.py`` (reference lines 8-171) on libgode.

Same class names, constructor arguments and ``state_dict`` keys (``fR.mlp.0.weight`` ... ``fR.output_linear.weight``,
``odefunc.fO.mlp.2.bias`` ...), so checkpoints of ``train_IN.py`` (``model/<name>.<fold>.pth``) load either way.

What changes is the arithmetic underneath.  The reference holds the relation structure as two DENSE one-hot matrices
``Msrc, Mtgt`` [n objects, m relations] and multiplies by them (model.py:65-66,71,113-114,118: 6 000 x 30 000 floats per
batch of 1 000 six-body systems); here each is turned ONCE into an index plan (``ops.plan_for``, cached on the tensor) and

    Msrc.t() @ O , Mtgt.t() @ O   are row gathers        (gode_spmm_csr_f32 on the transposed one-hot plan)
    Mtgt @ E                       is a segmented sum     (the same kernel; deterministic, no atomics)

while every ``nn.Linear`` (+ ReLU) of the two MLPs runs on ``ops.LinearFn`` (libgode GEMMs, ReLU in the epilogue).  The time
column of ``IN_ODEfunc`` stays a concatenated column: at 11 input features the product is negligible next to the gathers.
``Msrc`` / ``Mtgt`` may also be passed as ``ops.GraphPlan`` objects built from index lists (no dense matrix at all).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import odeint as _solver
from ... import ops


class _Linear(nn.Linear):
    """``nn.Linear`` parameters (weight [out, in]); forward on libgode with an optional fused ReLU."""

    def forward(self, x, relu=False):
        return ops.LinearFn.apply(x, self.weight.t(), self.bias, relu)


class MLP(nn.Module):
    """model.py:8-25: ``hidden_linear`` is a plain list there too (the layers are registered through ``self.mlp``), and
    ``output_linear`` is registered twice (attribute + last element of ``mlp``) -- both kept for the state_dict keys."""

    def __init__(self, input, layers, output):
        super().__init__()
        ds = [input] + layers + [output]
        self.hidden_linear = [_Linear(d_i, d_o, bias=True) for d_i, d_o in zip(ds[:-2], ds[1:-1])]
        self.output_linear = _Linear(ds[-2], ds[-1], bias=True)
        seq = []
        for l in self.hidden_linear:
            seq.append(l)
            seq.append(nn.ReLU())
        seq.append(self.output_linear)
        self.mlp = nn.Sequential(*seq)          # parameter container / key layout; forward() fuses Linear + ReLU

    def forward(self, x):
        for l in self.hidden_linear:
            x = l(x, relu=True)
        return self.output_linear(x)


# layer widths of the two learned functions (model.py:55-63,83-96,143-151: the same numbers in all three classes)
_RELATION_HIDDEN = 4 * [150]     # fR: relation model, four hidden layers
_EFFECT_WIDTH = 50               # d_E
_OBJECT_HIDDEN = 1 * [100]       # fO: object model, one hidden layer


def _learned_functions(relation_in, external_in, d_P):
    """(fR, fO) = (MLP(relation_in -> 150 x4 -> d_E), MLP(d_E + external_in -> 100 -> d_P))"""
    return (MLP(relation_in, list(_RELATION_HIDDEN), _EFFECT_WIDTH),
            MLP(_EFFECT_WIDTH + external_in, list(_OBJECT_HIDDEN), d_P))


def _plans(Msrc, Mtgt):
    return ops.plan_for(Msrc), ops.plan_for(Mtgt)


def _gather_T(plan, O):
    """``M.t() @ O`` for a planned one-hot M [n, m]: row r of the result is O[object of relation r]."""
    return ops.PlanMatmulFn.apply(O, plan.transposed(), None)


class IN(nn.Module):
    """model.py:29-74 (Battaglia et al.): E = fR([R, Msrc^T O, Mtgt^T O]); P = fO([Mtgt E, X])."""

    def __init__(self, d_O, d_R, d_X, d_P):
        super().__init__()
        self.fR, self.fO = _learned_functions(d_R + 2 * d_O, d_X, d_P)

    def forward(self, O, R, X, Msrc, Mtgt):
        psrc, ptgt = _plans(Msrc, Mtgt)
        Rsrc = _gather_T(psrc, O)
        Rtgt = _gather_T(ptgt, O)
        R_prime = torch.cat(([R] if R is not None else []) + [Rsrc, Rtgt], dim=1)
        E = self.fR(R_prime)
        E_prime = torch.cat([ops.PlanMatmulFn.apply(E, ptgt, None)] + ([X] if X is not None else []), dim=1)
        return self.fO(E_prime)


class IN_ODEfunc(nn.Module):
    """model.py:77-120: dx/dt = fO(Mtgt fR([Msrc^T O, Mtgt^T O, t])) with O = [x, fixed tail]; relation and external
    inputs are not supported there either (d_X = d_R = 0)."""

    def __init__(self, d_O, d_R, d_X, d_P):
        super().__init__()
        # the ODE function takes neither relation nor external inputs (d_R = d_X = 0 whatever is passed); + 1 = the time column
        self.fR, self.fO = _learned_functions(2 * d_O + 1, 0, d_P)
        self.nfe = 0

    def set_fixed(self, Otail, Msrc, Mtgt):
        self.Ofixed = Otail
        self.psrc, self.ptgt = _plans(Msrc, Mtgt)

    def forward(self, t, x):
        self.nfe += 1
        O = torch.cat([x, self.Ofixed], dim=1)
        Rsrc = _gather_T(self.psrc, O)
        Rtgt = _gather_T(self.ptgt, O)
        tt = torch.ones_like(Rsrc[:, :1]) * t
        E = self.fR(torch.cat([Rsrc, Rtgt, tt], dim=1))
        return self.fO(ops.PlanMatmulFn.apply(E, self.ptgt, None))


class IN_ODE(nn.Module):
    """model.py:123-171: ``odeint_adjoint(odefunc, O[:, :d_P], [0, 1], rtol = atol = tol)[-1]`` (dopri5).  ``fR`` / ``fO`` at
    this level are never called by the reference either; they exist for the parameter set and the state_dict."""

    def __init__(self, d_O, d_R, d_X, d_P, tol=1e-5, method=None, options=None):
        super().__init__()
        self.fR, self.fO = _learned_functions(d_R + 2 * d_O, d_X, d_P)
        self.odefunc = IN_ODEfunc(d_O, d_R, d_X, d_P)
        self.integration_time = torch.tensor([0, 1]).float()
        self.tol = tol
        self.d_P = d_P
        self.method, self.options = method, options     # builder extension (None = dopri5, as the reference)
        self.stats = None

    def forward(self, O, R, X, Msrc, Mtgt, tol=None):
        Otail = O[:, self.d_P:]
        Ohead = O[:, :self.d_P]
        self.odefunc.set_fixed(Otail, Msrc, Mtgt)
        # P[-1] of the reference's stacked result (no [2, n, d_P] stack is materialised)
        return _solver.odeint_adjoint_final(self.odefunc, Ohead.contiguous(), self.integration_time, rtol=self.tol,
                                            atol=self.tol, method=self.method, options=self.options, stats=self.stats)

    @property
    def nfe(self):
        return self.odefunc.nfe

    @nfe.setter
    def nfe(self, value):
        self.odefunc.nfe = value
