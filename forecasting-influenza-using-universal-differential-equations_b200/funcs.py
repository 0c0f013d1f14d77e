"""ODE-func modules with the reference's protocol, and their recognition.

``Fp`` / ``Fa`` / ``FaFp`` mirror the public surface of
/root/reference/lib/models.py:109-265 — constructor kwargs, ``forward(t, x)``,
``clear_tracking()``, ``posterior()``, attributes ``params``, ``tracker``,
``Fa_w``, ``ode_type``, ``uncertainty``, ``n_regions``, ``latent_dim`` and the
``state_dict`` keys (``Fp_net.{1,3,..}``, ``net.*``/``aug_net.*``) — so that a
checkpoint written by the reference loads here and vice versa.

The fused solver never calls ``forward``: ``describe_func`` reads the layer
structure and hands the flat weight vector to the CUDA kernels.  ``forward`` is
kept because it is part of the ODE-func protocol (``func(t, y)``) that callers
such as plotting code use directly.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn
from torch.distributions import Normal

from . import _lib


def _mlp_modules(n_in: int, sizes: Sequence[int], n_head: int, flatten_first: bool) -> nn.ModuleList:
    """[Flatten?], Linear, (ELU, Linear)*, Linear — index layout as in the reference
    (lib/models.py:116-124), which is what fixes the state_dict keys."""
    mods = nn.ModuleList()
    if flatten_first:
        mods.append(nn.Flatten())
    widths = [n_in, *sizes]
    mods.append(nn.Linear(widths[0], widths[1]))
    for a, b in zip(widths[1:-1], widths[2:]):
        mods.append(nn.ELU(inplace=True))
        mods.append(nn.Linear(a, b))
    mods.append(nn.Linear(widths[-1], n_head))
    return mods


def _run(mods: nn.ModuleList, x: torch.Tensor) -> torch.Tensor:
    for m in mods:
        x = m(x)
    return x


class Decoder(nn.Module):
    """The consumer of the solve's outputs, mirror of /root/reference/lib/models.py:26-51: ``Sequential(Flatten,
    Linear(R * latent_dim, R * input_dim))`` over ``data[..., :latent_dim]`` with ``latent_dim = 3`` when ``Fp=True`` (the
    reference's setting), N(0, 0.1) weights and zero bias (lib/utils.py:69-73).  Same ``state_dict`` keys (``decoder.1.*``).
    ``ude_b200.odeint_decoded`` applies it inside the solve's output store."""

    def __init__(self, n_regions, latent_dim, input_dim, Fp=True, device=torch.device("cpu"), dtype=torch.float32, **kwargs):
        super().__init__()
        self.Fp = Fp
        self.n_regions = n_regions
        self.input_dim = input_dim
        self.latent_dim = 3 if Fp else latent_dim
        self.decoder = nn.Sequential(nn.Flatten(), nn.Linear(n_regions * self.latent_dim, n_regions * input_dim)).to(device=device, dtype=dtype)
        nn.init.normal_(self.decoder[1].weight, mean=0, std=0.1)
        nn.init.constant_(self.decoder[1].bias, val=0)

    def forward(self, data):
        data = data[..., :self.latent_dim]
        shape = data.shape
        out = self.decoder(data.reshape((-1,) + tuple(shape[2:])))
        return out.reshape(tuple(shape[:2]) + (-1,))


class _UdeFunc(nn.Module):
    """Shared behaviour of the three ODE funcs."""
    ode_type = "?"

    def __init__(self, n_regions: int, latent_dim: int):
        super().__init__()
        self.n_regions = n_regions
        self.latent_dim = latent_dim
        self.uncertainty = "none"
        self.params: list = []
        self.tracker: list = []
        self._side_sums: list = []      # fused solves: the raw side sums behind the entries of `params` (one per solve)

    # -- reference protocol -----------------------------------------------------------
    def clear_tracking(self):
        self.params = []
        self.tracker = []
        self._side_sums = []

    def posterior(self) -> Normal:
        """Normal(mean, unbiased std) over every (beta, gamma) row since clear_tracking() (lib/models.py:152-156).  After fused
        solves the rows exist only as in-kernel sums: ONE solve is served by the surrogate entry in `params` (exactly what
        the reference's expression returns on it); SEVERAL solves are pooled exactly from the raw sums (their surrogate rows
        would weight every solve equally, whatever its row count)."""
        sums, self._side_sums = self._side_sums, []
        if len(sums) > 1 and len(sums) == len(self.params):
            from .odeint import posterior_moments
            self.params = []
            mean, std = posterior_moments(torch.stack(sums).sum(0))
            return Normal(mean.to(torch.float32), std.to(torch.float32), validate_args=False)
        rows = torch.stack(self.params).reshape(-1, 2)
        self.params = []
        # Argument validation reads a device flag back to the host (a sync, and illegal under CUDA-graph capture);
        # off while a stream is capturing -- same distribution either way.
        capturing = rows.is_cuda and torch.cuda.is_current_stream_capturing()
        return Normal(rows.mean(0), rows.std(0), validate_args=False if capturing else None)

    # -- shared math ------------------------------------------------------------------
    def _rates(self, net: nn.ModuleList, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        p = torch.abs(_run(net, x)).reshape(-1, self.n_regions, 2)
        self.params.append(p)
        infect = p[..., 0] * x[..., 0] * x[..., 1]
        recover = p[..., 1] * x[..., 1]
        return torch.stack([-infect, infect - recover, recover], dim=-1), p

    def _finish(self, dyn: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        out = torch.cat([dyn, torch.zeros_like(x[..., 3:])], -1)
        out[(x > 2) | (x < -1)] = 0.0
        return out


class Fp(_UdeFunc):
    """Mechanistic SIR term with MLP-produced (beta, gamma) — lib/models.py:109-156."""

    def __init__(self, n_regions=1, latent_dim=8, net_sizes=[20, 20], **kwargs):
        super().__init__(n_regions, latent_dim)
        self.ode_type = "Fp"
        self.Fp_net = _mlp_modules(n_regions * latent_dim, net_sizes, 2 * n_regions, flatten_first=True)

    def forward(self, t, x):
        dyn, _ = self._rates(self.Fp_net, x)
        return self._finish(dyn, x)


class Fa(_UdeFunc):
    """Pure MLP correction — lib/models.py:158-197 (Flatten is a separate attribute there)."""

    def __init__(self, n_regions=1, latent_dim=8, net_sizes=[32, 32], aug_net_sizes=[32, 32], nhidden_fa=32, **kwargs):
        super().__init__(n_regions, latent_dim)
        self.ode_type = "Fa"
        self.flatten = nn.Flatten()
        self.aug_net = _mlp_modules(n_regions * latent_dim, aug_net_sizes, 3 * n_regions, flatten_first=False)

    def forward(self, t, x):
        fa = _run(self.aug_net, self.flatten(x)).reshape(-1, self.n_regions, 3)
        out = self._finish(fa, x)
        self.tracker.append(fa)
        return out


class FaFp(_UdeFunc):
    """SIR term + Fa_w * correction — lib/models.py:199-265."""

    def __init__(self, n_regions=1, latent_dim=8, net_sizes=[20, 20], aug_net_sizes=[32, 32], **kwargs):
        super().__init__(n_regions, latent_dim)
        self.ode_type = "FaFp"
        self.net = _mlp_modules(n_regions * latent_dim, net_sizes, 2 * n_regions, flatten_first=True)
        self.aug_net = _mlp_modules(n_regions * latent_dim, aug_net_sizes, 3 * n_regions, flatten_first=True)
        self.Fa_w = 1.0

    def forward(self, t, x):
        dyn, _ = self._rates(self.net, x)
        fa = _run(self.aug_net, x).reshape(-1, self.n_regions, 3)
        out = self._finish(dyn + self.Fa_w * fa, x)
        self.tracker.append(fa)
        return out


# ---------------------------------------------------------------------------------------
# Bayesian ODE funcs -- mirror of /root/reference/lib/in_development/models_bayes.py
# ---------------------------------------------------------------------------------------
class Dense_Variational(nn.Module):
    """Linear layer with a factorised Gaussian over its weights, re-sampled at EVERY call:
    ``w = w_mean + z * |w_std|`` (models_bayes.py:12-48).  Same parameter names (``w_mean, w_std, b_mean, b_std``: the
    ``state_dict`` keys), same initialisation (:34-42), ``make_prior`` / ``make_posterior`` for the KL term (:50-62)."""

    def __init__(self, in_features, out_features, bias=True, prior_std=1.0):
        super().__init__()
        if not bias:
            raise NotImplementedError("the reference only builds Dense_Variational layers with a bias")
        self.in_features, self.out_features, self.prior_std, self.bias = in_features, out_features, prior_std, True
        self.w_mean = nn.Parameter(torch.empty(out_features, in_features))
        self.w_std = nn.Parameter(torch.empty(out_features, in_features))
        self.b_mean = nn.Parameter(torch.empty(out_features))
        self.b_std = nn.Parameter(torch.empty(out_features))
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.w_mean, a=5 ** 0.5)
        nn.init.constant_(self.w_std, 0.1)
        bound = 1 / self.in_features ** 0.5
        nn.init.uniform_(self.b_mean, -bound, bound)
        nn.init.constant_(self.b_std, 0.1)

    def make_z(self):
        with torch.no_grad():
            self.z = [torch.randn_like(self.w_mean), torch.randn_like(self.b_mean)]

    def forward(self, x):
        self.make_z()
        return nn.functional.linear(x, self.w_mean + self.z[0] * torch.abs(self.w_std), self.b_mean + self.z[1] * torch.abs(self.b_std))

    def make_prior(self):
        self.prior = [Normal(torch.zeros_like(self.w_mean), self.prior_std * torch.ones_like(self.w_mean)),
                      Normal(torch.zeros_like(self.b_mean), self.prior_std * torch.ones_like(self.b_mean))]
        return self.prior

    def make_posterior(self):
        self.posterior = [Normal(self.w_mean, torch.abs(self.w_std)), Normal(self.b_mean, torch.abs(self.b_std))]
        return self.posterior

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, bias=True"


def _var_modules(n_in: int, sizes: Sequence[int], n_head: int, prior_std: float) -> nn.ModuleList:
    """Flatten, DV, (ELU, DV)*, DV -- the index layout of models_bayes.py:78-84 (state_dict keys)."""
    mods = nn.ModuleList([nn.Flatten()])
    widths = [n_in, *sizes]
    mods.append(Dense_Variational(widths[0], widths[1], prior_std=prior_std))
    for a, b in zip(widths[1:-1], widths[2:]):
        mods.append(nn.ELU(inplace=True))
        mods.append(Dense_Variational(a, b, prior_std=prior_std))
    mods.append(Dense_Variational(widths[-1], n_head, prior_std=prior_std))
    return mods


class _BayesFunc(_UdeFunc):
    def __init__(self, n_regions, latent_dim):
        super().__init__(n_regions, latent_dim)
        self.uncertainty = "bayes"

    def get_kl(self):
        """Mean KL(posterior || prior) over the variational layers (models_bayes.py:118-128, 174-184, 254-265); differentiable
        torch ops on the parameters -- the fused solve does not touch it."""
        from torch.distributions import kl_divergence
        kl, count = 0, 0
        for net in self.children():
            for layer in net:
                if isinstance(layer, Dense_Variational):
                    kl = kl + sum(kl_divergence(q, p).mean() for q, p in zip(layer.make_posterior(), layer.make_prior())) / 2
                    count += 1
        return kl / count


class Bayes_Fp(_BayesFunc):
    """models_bayes.py:69-128."""
    ode_type = "Fp"

    def __init__(self, n_regions=1, latent_dim=8, net_sizes=(20, 20), prior_std=0.1, **kwargs):
        super().__init__(n_regions, latent_dim)
        self.Fp_net = _var_modules(n_regions * latent_dim, list(net_sizes), 2 * n_regions, prior_std)

    def forward(self, t, x):
        dyn, _ = self._rates(self.Fp_net, x)
        return self._finish(dyn, x)


class Bayes_Fa(_BayesFunc):
    """models_bayes.py:131-184."""
    ode_type = "Fa"

    def __init__(self, n_regions=1, latent_dim=8, aug_net_sizes=(32, 32), nhidden_fa=32, prior_std=0.1, **kwargs):
        super().__init__(n_regions, latent_dim)
        self.aug_net = _var_modules(n_regions * latent_dim, list(aug_net_sizes), 3 * n_regions, prior_std)

    def forward(self, t, x):
        fa = _run(self.aug_net, x).reshape(-1, self.n_regions, 3)
        out = self._finish(fa, x)
        self.tracker.append(fa)
        return out


class Bayes_FaFp(_BayesFunc):
    """models_bayes.py:186-265 (the rate net is called Fp_net here, not net as in lib/models.py FaFp)."""
    ode_type = "FaFp"

    def __init__(self, n_regions=1, latent_dim=8, net_sizes=(20, 20), aug_net_sizes=(32, 32), prior_std=0.1, **kwargs):
        super().__init__(n_regions, latent_dim)
        self.Fp_net = _var_modules(n_regions * latent_dim, list(net_sizes), 2 * n_regions, prior_std)
        self.aug_net = _var_modules(n_regions * latent_dim, list(aug_net_sizes), 3 * n_regions, prior_std)
        self.Fa_w = 1.0

    def forward(self, t, x):
        dyn, _ = self._rates(self.Fp_net, x)
        fa = _run(self.aug_net, x).reshape(-1, self.n_regions, 3)
        out = self._finish(dyn + self.Fa_w * fa, x)
        self.tracker.append(fa)
        return out


class _VarLayer:
    """A Dense_Variational layer (ours or the reference's) seen through nn.Linear's attribute names: weight / bias are the
    means; w_std / b_std ride along."""

    def __init__(self, m):
        self.module = m
        self.in_features, self.out_features = m.in_features, m.out_features
        self.weight, self.bias, self.w_std, self.b_std = m.w_mean, m.b_mean, m.w_std, m.b_std


def _is_variational(m) -> bool:
    return all(hasattr(m, a) for a in ("w_mean", "w_std", "b_mean", "b_std"))


# ---------------------------------------------------------------------------------------
# recognition of a reference-protocol module (ours, lib.models.*, or tuning/tune_node.py:63 Fp)
# ---------------------------------------------------------------------------------------
@dataclass
class FuncInfo:
    kind: str
    n_regions: int
    latent_dim: int
    rate_layers: List[nn.Linear]
    aug_layers: List[nn.Linear]
    fa_w: float
    rate_head_elu: bool = False      # ELU between the last hidden layer and the head (tune_node.py Fp)
    aug_head_elu: bool = False
    kernels: Optional[str] = None    # None: library picks; 'ffma' | 'tensor' pin the RK4 kernel family
    bayes: bool = False              # Dense_Variational layers: weights re-sampled at every RHS evaluation (rk4, FFMA engine)

    def desc(self, batch: int, n_groups: int = 1) -> "_lib.UdeDesc":
        d = _lib.UdeDesc()
        d.kind = _lib.KIND[self.kind]
        d.n_regions = self.n_regions
        d.latent_dim = self.latent_dim
        d.n_rate_layers = len(self.rate_layers)
        for i, l in enumerate(self.rate_layers):
            d.rate_sizes[i] = l.out_features
        d.n_aug_layers = len(self.aug_layers)
        for i, l in enumerate(self.aug_layers):
            d.aug_sizes[i] = l.out_features
        d.fa_w = float(self.fa_w)
        d.n_groups = n_groups
        d.flags = (1 if self.rate_head_elu else 0) | (2 if self.aug_head_elu else 0)
        if self.kernels is not None:
            if self.kernels not in ("ffma", "tensor"):
                raise ValueError("kernels must be 'ffma' or 'tensor'")
            d.flags |= 4 if self.kernels == "ffma" else 8
        d.batch = batch
        return d

    def linears(self) -> List[nn.Linear]:
        return [*self.rate_layers, *self.aug_layers]

    def theta(self) -> torch.Tensor:
        """Flat weight vector in the C-ABI layout; differentiable w.r.t. the module parameters."""
        return torch.cat([p.reshape(-1) for l in self.linears() for p in (l.weight, l.bias)])


class UnsupportedFunc(TypeError):
    pass


def _linear_chain(net, what: str):
    """Validate Linear, (ELU, Linear)*, [ELU,] Linear (an optional leading Flatten is skipped).
    Returns (linears, elu_before_head)."""
    mods = [m for m in net if not isinstance(m, nn.Flatten)]
    mods = [_VarLayer(m) if _is_variational(m) else m for m in mods]
    lin: List[nn.Linear] = []
    expect_linear = True
    for i, m in enumerate(mods):
        if isinstance(m, (nn.Linear, _VarLayer)):
            if m.bias is None:
                raise UnsupportedFunc(f"{what}: Linear without bias")
            lin.append(m)
            expect_linear = False
        elif isinstance(m, nn.ELU):
            if expect_linear or m.alpha != 1.0:
                raise UnsupportedFunc(f"{what}: unexpected ELU at position {i}")
            expect_linear = True
        else:
            raise UnsupportedFunc(f"{what}: unsupported layer {type(m).__name__}")
    if len(lin) < 2 or len(lin) > _lib.UDE_MAX_LAYERS:
        raise UnsupportedFunc(f"{what}: needs 2..{_lib.UDE_MAX_LAYERS} Linear layers")
    # the reference puts an ELU before every hidden Linear except the head: positions must be
    # Linear, ELU, Linear, ..., ELU, Linear, Linear
    n_elu = sum(isinstance(m, nn.ELU) for m in mods)
    if not isinstance(mods[-1], (nn.Linear, _VarLayer)):
        raise UnsupportedFunc(f"{what}: the last layer must be Linear")
    if len({type(l) for l in lin}) != 1:
        raise UnsupportedFunc(f"{what}: nn.Linear and Dense_Variational layers cannot be mixed")
    head_elu = isinstance(mods[-2], nn.ELU)                  # nn.Sequential style (tuning/tune_node.py:68-75)
    if n_elu != len(lin) - (1 if head_elu else 2):
        raise UnsupportedFunc(f"{what}: layer pattern is not Linear,(ELU,Linear)*,[ELU,]Linear")
    for a, b in zip(lin[:-1], lin[1:]):
        if a.out_features != b.in_features:
            raise UnsupportedFunc(f"{what}: width mismatch")
    return lin, head_elu


def describe_func(func) -> FuncInfo:
    """Duck-type an ODE func following the reference protocol; raise UnsupportedFunc otherwise."""
    kind = getattr(func, "ode_type", None)
    if kind is None and type(func).__name__ == "Fp" and hasattr(func, "net"):
        kind = "Fp"                                    # tuning/tune_node.py:63-104 local Fp
    if kind not in ("Fp", "Fa", "FaFp"):
        raise UnsupportedFunc(
            "the fused UDE solver integrates Fp / Fa / FaFp modules (lib/models.py protocol) only; "
            f"got {type(func).__name__}.  There is deliberately no generic-callable or CPU fallback.")
    unc = getattr(func, "uncertainty", "none")
    if unc not in ("none", "bayes"):
        raise UnsupportedFunc(f"unknown uncertainty mode {unc!r}")
    R = int(getattr(func, "n_regions", getattr(func, "n_region", 0)))
    L = int(func.latent_dim)
    rate: List[nn.Linear] = []
    aug: List[nn.Linear] = []
    r_elu = a_elu = False
    if kind == "Fp":
        rate, r_elu = _linear_chain(func.Fp_net if hasattr(func, "Fp_net") else func.net, "Fp net")
    elif kind == "FaFp":
        rate, r_elu = _linear_chain(func.Fp_net if unc == "bayes" else func.net, "FaFp.net")     # models_bayes.py:194 names it Fp_net
        aug, a_elu = _linear_chain(func.aug_net, "FaFp.aug_net")
    else:
        aug, a_elu = _linear_chain(func.aug_net, "Fa.aug_net")
    for lin, head in ((rate, 2 * R), (aug, 3 * R)):
        if lin and (lin[0].in_features != R * L or lin[-1].out_features != head):
            raise UnsupportedFunc("layer widths do not match n_regions * latent_dim / the head size")
    if L < 3:
        raise UnsupportedFunc("latent_dim must be >= 3 (S, I, R)")
    bayes = any(isinstance(l, _VarLayer) for l in (*rate, *aug))
    if bayes != (unc == "bayes"):
        raise UnsupportedFunc("uncertainty == 'bayes' goes with Dense_Variational layers (and only with them)")
    return FuncInfo(kind, R, L, rate, aug, float(getattr(func, "Fa_w", 1.0)), r_elu, a_elu, bayes=bayes)
