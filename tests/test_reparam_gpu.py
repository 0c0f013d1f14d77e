"""the producer of y0 -- reparam + 1e-5 (lib/models.py:16-24, lib/VAE.py:120,130) -- fused into the solve's
prologue, its backward into the adjoint's epilogue.  Checked against the reference's own op sequence in torch followed by
the plain fused odeint (bit-equal forward, gradients of mean / std / eps / theta) and against the oracle on the CPU."""
import pytest
import torch

from tests.util import RTOL_PARITY, flat_grad, module_from_spec, rel_err

pytestmark = pytest.mark.gpu


def _reparam(eps, std, mean, n_samples, batch):            # lib/models.py:16-24, verbatim op order, + 1e-5 (VAE.py:130)
    z = eps * std + mean
    z = torch.concat([torch.abs(z[..., :2]), (1 - torch.abs(z[..., :2]).sum(-1)).unsqueeze(-1), z[..., 2:]], -1)
    return z.reshape((n_samples * batch,) + z.shape[2:]) + 1e-5


@pytest.mark.parametrize("R,L,S,b,kernels", [(1, 8, 16, 6, "ffma"), (1, 8, 64, 5, "tensor"), (3, 4, 8, 7, "ffma"), (49, 8, 4, 5, "ffma")])
def test_fused_reparam_equals_reparam_then_odeint(R, L, S, b, kernels):
    import ude_b200
    from oracle import odeint_ref
    from oracle.ude_rhs import OracleFunc, UdeSpec, init_theta
    spec = UdeSpec("FaFp", R, L, (64, 64, 32), (64, 64))
    theta = init_theta(spec, seed=3)
    g = torch.Generator().manual_seed(R * 100 + L)
    mean = torch.tensor([0.9, 0.02] + [0.0] * (L - 3)).expand(b, R, L - 1) + 0.01 * torch.randn(b, R, L - 1, generator=g)
    std = (torch.tensor([0.1, 0.05] + [1.0] * (L - 3)) * 0.3).expand(b, R, L - 1).clone()
    eps = torch.randn(S, b, R, L - 1, generator=g)
    eps[0, 0, 0, 0] = -40.0                                   # a negative z0: the |.| branch and its sign in the backward
    t = torch.arange(5.)
    G = torch.randn(5, S * b, R, L, generator=g)

    outs = {}
    for mode in ("fused", "plain"):
        mod = module_from_spec(spec, theta, 1.0, "cuda")
        e, s_, m = (x.cuda().requires_grad_(True) for x in (eps, std, mean))
        if mode == "fused":
            y = ude_b200.odeint_reparam(mod, e, s_, m, t, options=dict(step_size=1.0, kernels=kernels))
        else:
            y = ude_b200.odeint(mod, _reparam(e, s_, m, S, b), t, method="rk4", options=dict(step_size=1.0, kernels=kernels))
        (y * G.cuda()).sum().backward()
        outs[mode] = (y.detach(), e.grad, s_.grad, m.grad, flat_grad(mod))
    assert torch.equal(outs["fused"][0], outs["plain"][0]), "the prologue must produce the bits of torch's reparam"
    for i, name in ((1, "eps"), (2, "std"), (3, "mean"), (4, "theta")):
        assert rel_err(outs["fused"][i], outs["plain"][i]) <= 2e-6, name

    f = OracleFunc(spec, theta)
    e, s_, m = (x.clone().requires_grad_(True) for x in (eps, std, mean))
    yr = odeint_ref.odeint(f, _reparam(e, s_, m, S, b), t, method="rk4", options=dict(step_size=1.0))
    (yr * G).sum().backward()
    assert rel_err(outs["fused"][0], yr) <= RTOL_PARITY
    assert rel_err(outs["fused"][2], s_.grad) <= RTOL_PARITY and rel_err(outs["fused"][3], m.grad) <= RTOL_PARITY
    assert rel_err(outs["fused"][4], f.theta.grad) <= RTOL_PARITY
