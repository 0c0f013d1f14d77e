"""the consumers of the solve's outputs -- Decoder Linear(3R -> R * input_dim) on latent[..., :3]
(lib/models.py:26-51, lib/VAE.py:138) and latent_init_loss(latent[..., :3]) (lib/train_functions.py:116-126, lib/VAE.py:186)
-- fused into the output store, their backward into the adjoint.  Checked against the CPU oracle followed by the
reference's own torch ops (values, and the gradients of y0 / theta / decoder weight / decoder bias), and against the plain
fused odeint followed by the same torch ops on the GPU."""
import pytest
import torch

from tests.util import RTOL_PARITY, flat_grad, module_from_spec, rel_err

pytestmark = pytest.mark.gpu


def latent_init_loss(x):                                   # lib/train_functions.py:116-126
    neg = torch.where(x < 0, abs(x), torch.zeros_like(x))
    big = torch.where(x > 1, abs(1 - x), torch.zeros_like(x))
    return (neg + big).sum()


def _loss(y_pred, reg, target, post=None, tracker=None):
    loss = ((y_pred - target) ** 2).mean() + 0.1 * reg / y_pred.shape[1]
    if post is not None:
        loss = loss + 0.01 * (post.mean ** 2).sum() + 0.01 * (post.stddev ** 2).sum()
    if tracker is not None:
        loss = loss + 0.01 * torch.norm(torch.stack(tracker)) / y_pred.shape[1] ** 0.5
    return loss


CASES = [
    # kind, R, L, n_dec, B, grid, kernels
    ("FaFp", 1, 8, 1, 300, "weekly", "ffma"),
    ("FaFp", 1, 8, 1, 300, "weekly", "tensor"),
    ("FaFp", 1, 8, 1, 200, "interp", "tensor"),
    ("FaFp", 1, 8, 1, 77, "interp", "ffma"),
    ("Fp", 1, 4, 2, 130, "daily", "ffma"),
    ("FaFp", 3, 8, 3, 100, "interp", "ffma"),
    ("FaFp", 10, 8, 10, 70, "weekly", "ffma"),
    ("Fa", 2, 16, 4, 90, "weekly", "ffma"),
    ("FaFp", 49, 8, 49, 40, "weekly", "ffma"),
]


def _grid(name):
    if name == "weekly":
        t = torch.arange(6.)
        return t, 1.0
    if name == "daily":
        t = torch.arange(15.) / 7
        return t, t[1] - t[0]
    t = torch.linspace(1, 21, 21) / 7          # tuning/tune_node.py:192,204: off-grid outputs, short last step
    return t, 1.0


@pytest.mark.parametrize("kind,R,L,n_dec,B,grid,kernels", CASES)
def test_decoded_solve_matches_oracle_plus_torch_consumers(kind, R, L, n_dec, B, grid, kernels):
    import ude_b200
    from oracle import odeint_ref
    from oracle.ude_rhs import OracleFunc, UdeSpec, init_theta, synth_y0
    spec = UdeSpec(kind, R, L, (64, 64, 32) if kind != "Fa" else (), (64, 64) if kind != "Fp" else ())
    theta = init_theta(spec, seed=11)
    g = torch.Generator().manual_seed(1000 + R)
    y0 = synth_y0(B, R, L, seed=5)
    y0[0, 0, 0] = 1.3            # S > 1 and R < 0: both branches of latent_init_loss (and the RHS clamp stays inactive)
    y0[0, 0, 2] = -0.32
    y0[1, 0, 1] = -0.01
    t, h = _grid(grid)
    dec_w = 0.3 * torch.randn(n_dec, 3 * R, generator=g)
    dec_b = 0.1 * torch.randn(n_dec, generator=g)
    target = torch.rand(t.numel(), 1, n_dec, generator=g)

    # ---- reference: CPU oracle + the reference's torch ops
    f = OracleFunc(spec, theta)
    y0r, wr, br = y0.clone().requires_grad_(True), dec_w.clone().requires_grad_(True), dec_b.clone().requires_grad_(True)
    lat = odeint_ref.odeint(f, y0r, t, method="rk4", options=dict(step_size=h))
    x3 = lat[..., :3]
    ypr = torch.nn.functional.linear(x3.reshape(t.numel(), B, 3 * R), wr, br)
    regr = latent_init_loss(x3)
    post = f.posterior() if spec.has_rate else None
    _loss(ypr, regr, target, post, f.tracker if spec.has_aug else None).backward()

    # ---- fused
    mod = module_from_spec(spec, theta, 1.0, "cuda")
    lin = torch.nn.Linear(3 * R, n_dec).cuda()
    with torch.no_grad():
        lin.weight.copy_(dec_w)
        lin.bias.copy_(dec_b)
    y0g = y0.cuda().requires_grad_(True)
    mod.clear_tracking()
    yp, reg = ude_b200.odeint_decoded(mod, lin, y0g, t, options=dict(step_size=h, kernels=kernels))
    assert yp.shape == (t.numel(), B, n_dec)
    post_g = mod.posterior() if spec.has_rate else None
    _loss(yp, reg, target.cuda(), post_g, mod.tracker if spec.has_aug else None).backward()

    assert float(regr) > 0.3
    assert rel_err(yp, ypr) <= RTOL_PARITY
    assert abs(float(reg) - float(regr)) <= 1e-5 * float(regr)
    assert rel_err(y0g.grad, y0r.grad) <= RTOL_PARITY
    assert rel_err(flat_grad(mod), f.theta.grad) <= RTOL_PARITY
    assert rel_err(lin.weight.grad, wr.grad) <= RTOL_PARITY
    assert rel_err(lin.bias.grad, br.grad) <= RTOL_PARITY


@pytest.mark.parametrize("kernels,B", [("ffma", 500), ("tensor", 148 * 128 + 50)])
def test_decoded_equals_plain_solve_plus_torch_ops_and_reference_decoder(kernels, B):
    """Same kernels with and without the fused epilogue: bit-equal reg mask semantics, y_pred to fp32 round-off; the
    decoder may be the package's Decoder mirror (or the reference's own class: same attribute layout)."""
    import ude_b200
    from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
    spec = UdeSpec("FaFp", 1, 8, (64, 64, 32), (64, 64))
    theta = init_theta(spec, seed=2)
    y0 = synth_y0(B, 1, 8, seed=9).cuda()
    t = torch.arange(9.)
    torch.manual_seed(0)
    dec = ude_b200.Decoder(1, 8, 1).cuda()
    target = torch.rand(9, 1, 1).cuda()
    res = {}
    for mode in ("fused", "plain"):
        mod = module_from_spec(spec, theta, 1.0, "cuda")
        dec.zero_grad(set_to_none=True)
        y0g = y0.clone().requires_grad_(True)
        if mode == "fused":
            yp, reg = ude_b200.odeint_decoded(mod, dec, y0g, t, options=dict(step_size=1.0, kernels=kernels))
        else:
            lat = ude_b200.odeint(mod, y0g, t, method="rk4", options=dict(step_size=1.0, kernels=kernels))
            yp = dec(lat[..., :3])
            reg = latent_init_loss(lat[..., :3])
        _loss(yp, reg, target).backward()
        res[mode] = (yp.detach(), float(reg), y0g.grad, flat_grad(mod), dec.decoder[1].weight.grad.clone(), dec.decoder[1].bias.grad.clone())
    assert rel_err(res["fused"][0], res["plain"][0]) <= 1e-6
    assert abs(res["fused"][1] - res["plain"][1]) <= 1e-6 * max(1.0, abs(res["plain"][1]))
    for i in (2, 3, 4, 5):
        assert rel_err(res["fused"][i], res["plain"][i]) <= 3e-6, i


def test_decoded_with_fused_reparam():
    """8(f1) + 8(f2) together: (eps, std, mean) in, (y_pred, reg) out, gradients of mean / std / decoder / theta."""
    import ude_b200
    from oracle.ude_rhs import UdeSpec, init_theta
    spec = UdeSpec("FaFp", 1, 8, (64, 64, 32), (64, 64))
    theta = init_theta(spec, seed=4)
    S, b, L = 32, 6, 8
    g = torch.Generator().manual_seed(7)
    mean = (torch.tensor([0.9, 0.02] + [0.0] * (L - 3)).expand(b, 1, L - 1) + 0.01 * torch.randn(b, 1, L - 1, generator=g)).cuda()
    std = (torch.tensor([0.1, 0.05] + [1.0] * (L - 3)) * 0.3).expand(b, 1, L - 1).clone().cuda()
    eps = torch.randn(S, b, 1, L - 1, generator=g).cuda()
    t = torch.arange(5.)
    dec = ude_b200.Decoder(1, 8, 1).cuda()
    target = torch.rand(5, 1, 1, generator=g).cuda()
    res = {}
    for mode in ("fused", "plain"):
        mod = module_from_spec(spec, theta, 1.0, "cuda")
        dec.zero_grad(set_to_none=True)
        s_, m = std.clone().requires_grad_(True), mean.clone().requires_grad_(True)
        if mode == "fused":
            yp, reg = ude_b200.odeint_decoded(mod, dec, None, t, options=dict(step_size=1.0, kernels="ffma"), reparam=(eps, s_, m))
        else:
            lat = ude_b200.odeint_reparam(mod, eps, s_, m, t, options=dict(step_size=1.0, kernels="ffma"))
            yp, reg = dec(lat[..., :3]), latent_init_loss(lat[..., :3])
        _loss(yp, reg, target).backward()
        res[mode] = (yp.detach(), s_.grad, m.grad, flat_grad(mod), dec.decoder[1].weight.grad.clone())
    assert rel_err(res["fused"][0], res["plain"][0]) <= 1e-6
    for i in (1, 2, 3, 4):
        assert rel_err(res["fused"][i], res["plain"][i]) <= 3e-6, i
