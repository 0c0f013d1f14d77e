"""Oracle restatement of the reference UDE right-hand sides (TEST INFRASTRUCTURE ONLY).

Functional form of ``Fp.forward`` (/root/reference/lib/models.py:129-146),
``Fa.forward`` (:177-188) and ``FaFp.forward`` (:230-254): the weights come in
as ONE flat vector ``theta`` laid out exactly like the C-ABI expects
(include/ude_b200.h): for every ``nn.Linear`` in ModuleList order, rate net
first then the augmentation net, ``W`` row-major ``[out, in]`` followed by
``b``.

Checked bit-for-bit against the real reference classes by
``tests/test_oracle.py::test_rhs_restatement_matches_real_reference_modules`` and, for the Bayesian variants,
``tests/test_bayes.py::test_oracle_matches_real_reference_bayes_classes`` (both run wherever /root/reference exists)
and through the committed fixtures in ``tests/golden`` (everywhere).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class UdeSpec:
    """Static description of one ODE func (mirrors UdeDesc in include/ude_b200.h)."""
    kind: str                       # 'Fp' | 'Fa' | 'FaFp'
    n_regions: int
    latent_dim: int
    net_sizes: Tuple[int, ...] = ()      # hidden widths of the rate net (beta, gamma head)
    aug_sizes: Tuple[int, ...] = ()      # hidden widths of the augmentation net (Fa head)
    head_elu: bool = False               # ELU also before the heads (tuning/tune_node.py:63-76 Fp)

    def __post_init__(self):
        assert self.kind in ("Fp", "Fa", "FaFp")
        assert self.latent_dim >= 3

    @property
    def has_rate(self) -> bool:
        return self.kind in ("Fp", "FaFp")

    @property
    def has_aug(self) -> bool:
        return self.kind in ("Fa", "FaFp")

    def layer_dims(self) -> List[Tuple[str, int, int]]:
        """[(net, n_in, n_out)] in theta order."""
        d = self.n_regions * self.latent_dim
        out = []
        if self.has_rate:
            widths = [d, *self.net_sizes, 2 * self.n_regions]
            out += [("rate", a, b) for a, b in zip(widths[:-1], widths[1:])]
        if self.has_aug:
            widths = [d, *self.aug_sizes, 3 * self.n_regions]
            out += [("aug", a, b) for a, b in zip(widths[:-1], widths[1:])]
        return out

    def n_theta(self) -> int:
        return sum(i * o + o for _, i, o in self.layer_dims())

    def macs_per_eval(self) -> int:
        return sum(i * o for _, i, o in self.layer_dims())


def split_theta(spec: UdeSpec, theta: torch.Tensor):
    """theta -> {'rate': [(W,b)...], 'aug': [(W,b)...]} views."""
    nets = {"rate": [], "aug": []}
    off = 0
    for net, n_in, n_out in spec.layer_dims():
        W = theta[off:off + n_in * n_out].reshape(n_out, n_in)
        off += n_in * n_out
        b = theta[off:off + n_out]
        off += n_out
        nets[net].append((W, b))
    assert off == theta.numel(), (off, theta.numel())
    return nets


def _mlp(layers, u: torch.Tensor, head_elu: bool = False) -> torch.Tensor:
    """Linear -> [ELU -> Linear]* -> Linear : NO ELU before the head
    (/root/reference/lib/models.py:118-124) unless head_elu (tuning/tune_node.py:68-75)."""
    h = F.linear(u, *layers[0])
    for W, b in layers[1:-1]:
        h = F.linear(F.elu(h), W, b)
    if head_elu:
        h = F.elu(h)
    return F.linear(h, *layers[-1])


def ude_rhs(spec: UdeSpec, theta: torch.Tensor, x: torch.Tensor, fa_w: float = 1.0):
    """One RHS evaluation.  x: [B,R,L] -> (f [B,R,L], p [B,R,2] | None, fa [B,R,3] | None).

    p (= |head| = beta,gamma) and fa are the side-channel tensors the reference
    appends to ``self.params`` / ``self.tracker`` (pre-mask, un-weighted).
    """
    B, R, L = x.shape
    assert (R, L) == (spec.n_regions, spec.latent_dim)
    nets = split_theta(spec, theta)
    u = x.reshape(B, R * L)
    mask = (x > 2) | (x < -1)                                   # models.py:130
    dyn = torch.zeros(B, R, 3, dtype=x.dtype, device=x.device)
    p = fa = None
    if spec.has_rate:
        p = torch.abs(_mlp(nets["rate"], u, spec.head_elu)).reshape(B, R, 2)   # models.py:135
        plus_i = p[..., 0] * x[..., 0] * x[..., 1]              # (beta*S)*I, left assoc. (:139)
        minus_i = p[..., 1] * x[..., 1]
        dyn = torch.stack([-plus_i, plus_i - minus_i, minus_i], dim=-1)
    if spec.has_aug:
        fa = _mlp(nets["aug"], u, spec.head_elu).reshape(B, R, 3)
        if spec.kind == "Fa":
            dyn = fa                                            # models.py:184 (no Fa_w)
        else:
            dyn = dyn + fa_w * fa                               # models.py:249
    res = torch.cat([dyn, torch.zeros_like(x[..., 3:])], dim=-1)
    res = torch.where(mask, torch.zeros_like(res), res)        # res[mask] = 0 (:145/251)
    return res, p, fa


class OracleFunc(torch.nn.Module):
    """``f(t, y)`` callable with the reference's side-channel lists, driven by
    oracle.odeint_ref.  Holds theta as ONE leaf so dL/dtheta comes back flat."""

    def __init__(self, spec: UdeSpec, theta: torch.Tensor, fa_w: float = 1.0):
        super().__init__()
        self.spec = spec
        self.theta = torch.nn.Parameter(theta.detach().clone())
        self.Fa_w = fa_w
        self.params: list = []
        self.tracker: list = []
        self.n_evals = 0

    def forward(self, t, x):
        f, p, fa = ude_rhs(self.spec, self.theta, x, self.Fa_w)
        self.n_evals += 1
        if p is not None:
            self.params.append(p)
        if fa is not None:
            self.tracker.append(fa)
        return f

    def clear_tracking(self):
        self.params, self.tracker = [], []

    def posterior(self):
        """Normal(mean, unbiased std) over every (eval, trajectory, region) row of (beta, gamma);
        clears the list (/root/reference/lib/models.py:152-156, 261-265)."""
        p = torch.stack(self.params).reshape(-1, 2)
        self.params = []
        return torch.distributions.Normal(p.mean(0), p.std(0))

    def side_sums(self) -> torch.Tensor:
        """[sum beta, sum gamma, sum beta^2, sum gamma^2, sum Fa^2, N] in fp64 —
        the reductions the fused kernels emit instead of the python lists."""
        out = torch.zeros(6, dtype=torch.float64)
        if self.params:
            p = torch.cat([q.reshape(-1, 2) for q in self.params]).double()
            out[0:2] = p.sum(0)
            out[2:4] = (p * p).sum(0)
            out[5] = p.shape[0]
        if self.tracker:
            fa = torch.cat([q.reshape(-1, 3) for q in self.tracker]).double()
            out[4] = (fa * fa).sum()
            if not self.params:
                out[5] = fa.numel() // 3
        return out


class BayesOracleFunc(OracleFunc):
    """The Bayes_Fp / Bayes_Fa / Bayes_FaFp funcs of /root/reference/lib/in_development/models_bayes.py:69-265: the same
    RHS over Dense_Variational layers, whose weights are re-drawn at EVERY call (:30-48):
        w = w_mean + z * |w_std|        (z ~ N(0,1) per call; here row ``n_evals`` of the given ``z`` [n_calls, n_theta])
    ``mean`` / ``std`` are flat in theta order (W then b per layer).  Pinned against the real classes (their ``make_z``
    replaced by a reader of the same z) in tests/test_bayes.py."""

    def __init__(self, spec: UdeSpec, mean: torch.Tensor, std: torch.Tensor, z: torch.Tensor, fa_w: float = 1.0):
        super().__init__(spec, mean, fa_w)
        self.std = torch.nn.Parameter(std.detach().clone())
        self.z = z.detach().clone()

    def forward(self, t, x):
        theta_e = self.theta + self.z[self.n_evals] * torch.abs(self.std)
        f, p, fa = ude_rhs(self.spec, theta_e, x, self.Fa_w)
        self.n_evals += 1
        if p is not None:
            self.params.append(p)
        if fa is not None:
            self.tracker.append(fa)
        return f


def theta_from_module(mod: torch.nn.Module) -> Tuple[UdeSpec, torch.Tensor, float]:
    """Read (spec, flat theta, Fa_w) off a reference-protocol module
    (lib.models.Fp / Fa / FaFp or anything with the same attributes)."""
    kind = mod.ode_type
    R = int(mod.n_regions)
    L = int(mod.latent_dim)

    def linears(net):
        return [m for m in net if isinstance(m, torch.nn.Linear)]

    rate = linears(mod.Fp_net) if kind == "Fp" else (linears(mod.net) if kind == "FaFp" else [])
    aug = linears(mod.aug_net) if kind in ("Fa", "FaFp") else []
    spec = UdeSpec(kind, R, L,
                   tuple(l.out_features for l in rate[:-1]),
                   tuple(l.out_features for l in aug[:-1]))
    flat = torch.cat([t.detach().reshape(-1) for l in (*rate, *aug) for t in (l.weight, l.bias)])
    assert flat.numel() == spec.n_theta()
    return spec, flat, float(getattr(mod, "Fa_w", 1.0))


AUG_HEAD_SCALE = 0.05


def init_theta(spec: UdeSpec, seed: int = 0, dtype=torch.float32,
               aug_head_scale: float = AUG_HEAD_SCALE) -> torch.Tensor:
    """nn.Linear default init (the reference leaves the ODE nets at it,
    /root/reference/lib/models.py:118-124), laid out flat.

    ``aug_head_scale`` shrinks the augmentation head: with the raw init the
    correction term drives S past the clamp at 2 within a few weeks (a
    discontinuous RHS no parity tolerance survives); a trained model keeps Fa
    small through the ``Fa_norm`` penalty (/root/reference/lib/VAE.py:179-183).
    """
    g = torch.Generator().manual_seed(seed)
    parts = []
    dims = spec.layer_dims()
    for idx, (net, n_in, n_out) in enumerate(dims):
        bound = 1.0 / (n_in ** 0.5)
        is_aug_head = net == "aug" and (idx + 1 == len(dims))
        s = aug_head_scale if is_aug_head else 1.0
        parts.append((torch.rand(n_out * n_in, generator=g, dtype=torch.float64) * 2 - 1) * bound * s)
        parts.append((torch.rand(n_out, generator=g, dtype=torch.float64) * 2 - 1) * bound * s)
    return torch.cat(parts).to(dtype)


def synth_y0(B: int, R: int, L: int, seed: int = 0, dtype=torch.float32) -> torch.Tensor:
    """Synthetic ILI-shaped initial conditions: mirrors
    ``reparam`` (/root/reference/lib/models.py:16-24) with encoder-scale stds and
    the ``+1e-5`` of /root/reference/lib/VAE.py:130."""
    g = torch.Generator().manual_seed(seed)
    e = torch.randn(B, R, L, generator=g, dtype=torch.float64)
    S = (0.9 + 0.05 * e[..., 0]).abs()
    I = (0.02 + 0.01 * e[..., 1]).abs()
    y = torch.cat([S[..., None], I[..., None], (1 - S - I)[..., None], e[..., 3:]], -1) + 1e-5
    return y.to(dtype)
