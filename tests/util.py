"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import glob
import os

import numpy as np
import torch

from oracle.ude_rhs import UdeSpec

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Parity tolerance of BASELINE.json's north_star: max relative trajectory and gradient
# error <= 1e-5, measured as max|a-b| / max|b| over each tensor.
RTOL_PARITY = 1e-5


def golden_names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["spec"] = UdeSpec(str(g["kind"]), int(g["n_regions"]), int(g["latent_dim"]),
                        tuple(int(v) for v in g["net_sizes"]), tuple(int(v) for v in g["aug_sizes"]))
    g["fa_w"] = float(g["fa_w"])
    g["method"] = str(g["method"])
    return g


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| (b = reference)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = float(b.abs().max())
    if denom == 0.0:
        return float((a - b).abs().max())
    return float((a - b).abs().max()) / denom


def scaled_err(a: torch.Tensor, b: torch.Tensor, rtol: float = 1e-5, atol_frac: float = 1e-7) -> float:
    """Element-wise scaled error: max over elements of |a-b| / (rtol*|b| + atol_frac*max|b|) (b = reference).
    <= 1 means every element is inside its own tolerance -- unlike rel_err, a small component (I ~ 0.02 next to
    S ~ 0.9, a small dtheta entry) is judged at its OWN magnitude plus an fp32-noise floor tied to the tensor's scale.
    Trajectories use atol_frac = 1e-7 (~ one fp32 ulp of the largest entry); gradients are fp32 SUMS over
    trajectories x stages, whose round-off floor is a few 1e-7 of the largest entry in ANY fp32 implementation
    (the oracle's own autograd included), so gradient tests state a wider atol_frac explicitly."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    scale = rtol * b.abs() + atol_frac * float(b.abs().max())
    if float(scale.max()) == 0.0:
        return float((a - b).abs().max())
    return float(((a - b).abs() / scale.clamp_min(1e-300)).max())


def module_from_spec(spec: UdeSpec, theta: torch.Tensor, fa_w: float = 1.0, device="cpu"):
    """Build this package's Fp/Fa/FaFp with the weights of a flat theta."""
    import ude_b200
    kw = {}
    if spec.has_rate:
        kw["net_sizes"] = list(spec.net_sizes)
    if spec.has_aug:
        kw["aug_net_sizes"] = list(spec.aug_sizes)
    mod = getattr(ude_b200, spec.kind)(spec.n_regions, spec.latent_dim, **kw)
    info = ude_b200.describe_func(mod)
    off = 0
    with torch.no_grad():
        for lin in info.linears():
            n = lin.weight.numel()
            lin.weight.copy_(theta[off:off + n].reshape(lin.weight.shape))
            off += n
            n = lin.bias.numel()
            lin.bias.copy_(theta[off:off + n])
            off += n
    assert off == theta.numel()
    if spec.kind == "FaFp":
        mod.Fa_w = fa_w
    return mod.to(device)


def flat_grad(mod) -> torch.Tensor:
    import ude_b200
    info = ude_b200.describe_func(mod)
    return torch.cat([p.grad.reshape(-1) for l in info.linears() for p in (l.weight, l.bias)])
