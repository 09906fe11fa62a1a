"""GCN models with the reference's names, constructor signatures, ``forward(x, *adj)`` and ``state_dict``
keys (GCN/models.py:8-600), running on libgode kernels.

The reference file spells out ~20 near-identical classes; here one ``_Stack`` describes a model as
``input layer -> middle blocks -> output layer`` and the named classes only choose the blocks.
``ODEBlock`` calls the package's own solver (``..odeint``) where the reference calls torchdiffeq.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import odeint as _solver
from .. import ops
from .layers import FixedGraphConvolution, GraphConvolution


class GroupNorm(nn.GroupNorm):
    """nn.GroupNorm whose [N, d] forward/backward run on gode_groupnorm_fwd/bwd (same parameters/keys)."""

    def forward(self, x):
        if x.dim() == 2 and x.is_cuda:
            return ops.group_norm(x, self.num_groups, self.weight, self.bias, self.eps)
        raise TypeError("GroupNorm here expects a CUDA [N, d] tensor; graph-odenet_b200 has no CPU path")


def _norm(nhid):
    return GroupNorm(min(32, nhid), nhid)


# ---------------------------------------------------------------------------------------------
# ODE function / block (GCN/models.py:161-201, 551-575)
# ---------------------------------------------------------------------------------------------


class ODEfunc(nn.Module):
    """f(t, x) = relu(FixedGC([t*1 || GroupNorm(x)]))  (GCN/models.py:161-179)."""

    def __init__(self, dim):
        super().__init__()
        self.norm1 = _norm(dim)
        self.gc1 = FixedGraphConvolution(dim + 1, dim)
        self.nfe = 0

    def set_adj(self, adj):
        self.gc1.set_adj(adj)

    def _gode_fused(self):
        """Graph plan if the fused libgode ODE-function kernels apply to this function, else None."""
        adj = self.gc1.adj
        if isinstance(adj, ops.GraphPlan) or hasattr(adj, "make_kernel"):   # single-device or row-partitioned plan
            return adj
        if torch.is_tensor(adj) and adj.is_cuda and adj.dim() == 2 and adj.shape[0] == adj.shape[1] and adj.shape[0] > 1:
            return ops.plan_for(adj)
        return None

    def forward(self, t, x):
        # un-fused evaluation (used when the function is called directly, outside ODEBlock)
        self.nfe += 1
        xn = self.norm1(x)
        tt = torch.ones_like(xn[:, :1]) * t
        return self.gc1(torch.cat([tt, xn], 1), relu=True)


class ODEfunc2(nn.Module):
    """Two (FixedGC -> relu -> GroupNorm) stages, t prepended to each (GCN/models.py:551-575)."""

    def __init__(self, dim, dropout):
        super().__init__()
        self.norm1, self.norm2 = _norm(dim), _norm(dim)
        self.gc1 = FixedGraphConvolution(dim + 1, dim)
        self.gc2 = FixedGraphConvolution(dim + 1, dim)
        self.dropout = dropout
        self.nfe = 0

    def set_adj(self, adj):
        self.gc1.set_adj(adj)
        self.gc2.set_adj(adj)

    def forward(self, t, x):
        self.nfe += 1
        tt = torch.ones_like(x[:, :1]) * t
        x = self.norm1(self.gc1(torch.cat([tt, x], 1), relu=True))
        return self.norm2(self.gc2(torch.cat([tt, x], 1), relu=True))


class ODEBlock(nn.Module):
    """``odeint_adjoint(odefunc, x, [0, 1], rtol=atol=tol)[1]``  (GCN/models.py:181-201).

    ``method`` / ``options`` are builder extensions (the reference always runs torchdiffeq's default,
    dopri5): ``method='rk4'`` is the fixed-step 3/8-rule solver BASELINE.json's metric is quoted on.
    """

    def __init__(self, odefunc, tol=1e-5, method=None, options=None):
        super().__init__()
        self.odefunc = odefunc
        self.integration_time = torch.tensor([0, 1]).float()
        self.tol = tol
        self.method, self.options = method, options
        self.stats = None

    def forward(self, x, *adj):
        self.odefunc.set_adj(*adj)
        # integration_time stays on the host: the solver keeps time in float32 scalars, so there is no
        # device round trip per forward (the reference's .type_as(x) moves it to the GPU and reads it back)
        return _solver.odeint_adjoint_final(self.odefunc, x, self.integration_time, rtol=self.tol, atol=self.tol,
                                            method=self.method, options=self.options, stats=self.stats)

    @property
    def nfe(self):
        return self.odefunc.nfe

    @nfe.setter
    def nfe(self, value):
        self.odefunc.nfe = value


# ---------------------------------------------------------------------------------------------
# model skeleton
# ---------------------------------------------------------------------------------------------


class _NfeMixin:
    """``nfe`` plumbing + the layer family a model is built from (GCN here; GAT/models.py rebinds these).

    Every model computes its pre-softmax scores in ``logits(x, *adj)``; ``forward`` is the reference's
    ``F.log_softmax(..., dim=1)`` of them.  The trainer's fused loss (ops.log_softmax_nll) starts from ``logits``."""

    def forward(self, x, *adj):
        return F.log_softmax(self.logits(x, *adj), dim=1)

    _conv = GraphConvolution
    _odefunc = None     # bound below, after ODEfunc / ODEfunc2 are defined
    _odefunc2 = None

    @property
    def nfe(self):
        blocks = [m for m in self.modules() if isinstance(m, ODEBlock)]
        return blocks[0].nfe if len(blocks) == 1 else sum(b.nfe for b in blocks)

    @nfe.setter
    def nfe(self, value):
        for m in self.modules():
            if isinstance(m, ODEBlock):
                m.nfe = value


_NfeMixin._odefunc = ODEfunc
_NfeMixin._odefunc2 = ODEfunc2


def _drop(x, p, training):
    return F.dropout(x, p, training=training)


class GCN(_NfeMixin, nn.Module):
    """gc1 -> relu -> dropout -> gc2 -> log_softmax  (GCN/models.py:8-21)."""

    def __init__(self, nfeat, nhid, nclass, dropout):
        super().__init__()
        self.gc1 = self._conv(nfeat, nhid)
        self.gc2 = self._conv(nhid, nclass)
        self.dropout = dropout

    def logits(self, x, *adj):
        x = _drop(self.gc1(x, *adj, relu=True), self.dropout, self.training)
        return self.gc2(x, *adj)


class RGCN2(_NfeMixin, nn.Module):
    """GCN/models.py:23-40: residual on the output layer, first nclass columns are the logits."""

    def __init__(self, nfeat, nhid, nclass, dropout):
        super().__init__()
        if nhid < nclass:
            raise ValueError("nhid must be equal or larger than nclass")
        self.gc1 = self._conv(nfeat, nhid)
        self.gc2 = self._conv(nhid, nhid)
        self.nclass, self.dropout = nclass, dropout

    def logits(self, x, *adj):
        x = _drop(self.gc1(x, *adj, relu=True), self.dropout, self.training)
        x = self.gc2(x, *adj) + x
        return x[:, :self.nclass]


class ODEGCN2(_NfeMixin, nn.Module):
    """GCN/models.py:42-64.  (The reference forgets to store ``nclass`` and fails in forward; stored here.)"""

    def __init__(self, nfeat, nhid, nclass, dropout):
        super().__init__()
        if nhid < nclass:
            raise ValueError("nhid must be equal or larger than nclass")
        self.gc1 = self._conv(nfeat, nhid)
        self.gc2 = ODEBlock(self._odefunc(nhid))
        self.nclass, self.dropout = nclass, dropout

    def logits(self, x, *adj):
        x = self.gc2(self.gc1(x, *adj, relu=True), *adj)
        return x[:, :self.nclass]


class _Three(_NfeMixin, nn.Module):
    """gc1 -> [norm1] -> middle -> gc3 -> log_softmax; `middle` is what the named subclasses differ in."""

    residual = False      # add the middle block's input to its output
    mid_norm = False      # GroupNorm after the middle convolution instead of dropout
    in_norm = False       # GroupNorm after the input convolution instead of dropout
    ode = False

    def __init__(self, nfeat, nhid, nclass, dropout):
        super().__init__()
        self.gc1 = self._conv(nfeat, nhid)
        if self.in_norm:
            self.norm1 = _norm(nhid)
        self.gc2 = ODEBlock(self._odefunc(nhid)) if self.ode else self._conv(nhid, nhid)
        if self.mid_norm:
            self.norm2 = _norm(nhid)
        self.gc3 = self._conv(nhid, nclass)
        self.dropout = dropout

    def logits(self, x, *adj):
        x = self.gc1(x, *adj, relu=True)
        x = self.norm1(x) if self.in_norm else _drop(x, self.dropout, self.training)
        if self.ode:
            x = self.gc2(x, *adj)
        else:
            r = x
            x = self.gc2(x, *adj, relu=True)
            x = self.norm2(x) if self.mid_norm else _drop(x, self.dropout, self.training)
            if self.residual:
                x = x + r
        return self.gc3(x, *adj)


class GCN3(_Three):               # GCN/models.py:66-81
    pass


class GCN3norm(_Three):           # GCN/models.py:83-99
    mid_norm = True


class RGCN3(_Three):              # GCN/models.py:101-118
    residual = True


class RGCN3norm(_Three):          # GCN/models.py:120-138
    residual, mid_norm = True, True


class RGCN3fullnorm(_Three):      # GCN/models.py:140-159
    residual, mid_norm, in_norm = True, True, True


class ODEGCN3(_Three):            # GCN/models.py:204-226
    ode = True


class ODEGCN3fullnorm(_Three):    # GCN/models.py:229-253
    ode, in_norm = True, True


class _Deep(_NfeMixin, nn.Module):
    """K-layer models: ``gcs`` = [input conv, middle..., output conv] (+ ``norms``)  (GCN/models.py:255-600)."""

    min_layers = 2
    residue_layers = 0            # 0: plain stack; r >= 1: add a residual every r middle layers
    use_norm = False
    too_few = "Can't make a GCN with less than 2 layers"

    def __init__(self, nfeat, nhid, nclass, dropout, nlayers=None, residue_layers=None):
        super().__init__()
        nlayers = self.min_layers if nlayers is None else nlayers
        if residue_layers is not None:
            self.residue_layers = residue_layers
        if nlayers < self._min(self.residue_layers):
            raise ValueError(self._too_few_msg(self.residue_layers))
        self.n_layers = nlayers
        self.gcs = nn.ModuleList([self._conv(nfeat, nhid)] + self._middle(nhid, nlayers, dropout)
                                 + [self._conv(nhid, nclass)])
        if self.use_norm:
            self.norms = nn.ModuleList([_norm(nhid) for _ in range(nlayers - 2)])
        self.dropout = dropout

    def _min(self, r):
        return self.min_layers

    def _too_few_msg(self, r):
        return self.too_few

    def _middle(self, nhid, nlayers, dropout):
        return [self._conv(nhid, nhid) for _ in range(nlayers - 2)]

    def logits(self, x, *adj):
        x = _drop(self.gcs[0](x, *adj, relu=True), self.dropout, self.training)
        countdown, r = 1, None
        for i, gc in enumerate(self.gcs[1:-1]):
            if isinstance(gc, ODEBlock):
                x = gc(x, *adj)
                continue
            if self.residue_layers:
                countdown -= 1
                if countdown == 0:
                    r, countdown = x, self.residue_layers
            x = gc(x, *adj, relu=True)
            x = self.norms[i](x) if self.use_norm else _drop(x, self.dropout, self.training)
            if self.residue_layers and countdown == 1:
                x = x + r
        if self.residue_layers and countdown > 1:
            x = x + r
        return self.gcs[-1](x, *adj)


class GCNK(_Deep):                # GCN/models.py:255-278
    pass


class GCNKnorm(_Deep):            # GCN/models.py:280-306
    use_norm = True


class RESK1(_Deep):               # GCN/models.py:309-338
    min_layers, residue_layers = 3, 1
    too_few = "Can't make a Residual GCN with less than 3 layers using 1 layer for each residual block"


class RESK2(_Deep):               # GCN/models.py:341-375
    min_layers, residue_layers = 4, 2
    too_few = "Can't make a Residual GCN with less than 4 layers using 2 layers for each residual block"


class RESK(_Deep):                # GCN/models.py:378-412
    min_layers, residue_layers = 3, 1

    def _min(self, r):
        return 2 + r

    def _too_few_msg(self, r):
        return "Can't make a Residual GCN with less than {} layers using {} layers for each residual block".format(2 + r, r)


class RESK1norm(RESK1):           # GCN/models.py:415-446
    use_norm = True


class RESK2norm(RESK2):           # GCN/models.py:449-484
    use_norm = True


class RESKnorm(RESK):             # GCN/models.py:487-521
    use_norm = True


class ODEK1(_Deep):               # GCN/models.py:524-548
    min_layers = 3
    too_few = RESK1.too_few

    def _middle(self, nhid, nlayers, dropout):
        return [ODEBlock(self._odefunc(nhid)) for _ in range(nlayers - 2)]


class ODEK2(_Deep):               # GCN/models.py:578-600
    min_layers = 4
    too_few = RESK2.too_few

    def _middle(self, nhid, nlayers, dropout):
        # the reference passes `dropout` as the block's tolerance (models.py:587); kept for parity
        blocks = [ODEBlock(self._odefunc2(nhid, dropout), dropout) for _ in range((nlayers - 2) // 2)]
        if nlayers % 2 == 1:
            blocks.append(ODEBlock(self._odefunc(nhid)))
        return blocks


MODEL_NAMES = ("GCN", "RGCN2", "ODEGCN2", "GCN3", "GCN3norm", "RGCN3", "RGCN3norm", "RGCN3fullnorm", "ODEGCN3",
               "ODEGCN3fullnorm", "GCNK", "GCNKnorm", "RESK1", "RESK2", "RESK", "RESK1norm", "RESK2norm", "RESKnorm",
               "ODEK1", "ODEK2")


def rebind_family(namespace, conv, odefunc, odefunc2):
    """Define every model class of this file in ``namespace`` on another layer family (used by GAT/models.py,
    whose reference file is this one with ``adj`` replaced by ``src, tgt, Mtgt`` -- GAT/models.py:16-600)."""
    made = {}
    for name in MODEL_NAMES:
        made[name] = type(name, (globals()[name],), {"_conv": conv, "_odefunc": odefunc, "_odefunc2": odefunc2,
                                                     "__module__": namespace.get("__name__", __name__),
                                                     "__doc__": globals()[name].__doc__})
    namespace.update(made)
    return made
