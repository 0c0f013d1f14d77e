"""one training iteration (odeint -> loss -> backward -> gated Adam step) captured in a CUDA graph and an
epoch driven without a host round trip per iteration (lib/VAE.py:200-223, 248-291).  The captured iteration must do
exactly what the eager one does."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(seed=0, B=2048):
    import ude_b200
    from oracle.ude_rhs import synth_y0
    torch.manual_seed(seed)
    mod = ude_b200.FaFp(1, 8, net_sizes=[64, 64, 32], aug_net_sizes=[64, 64])
    with torch.no_grad():
        mod.aug_net[-1].weight.mul_(0.05)
        mod.aug_net[-1].bias.mul_(0.05)
    mod = mod.cuda()
    y0s = [synth_y0(B, 1, 8, seed=10 + i).cuda() for i in range(4)]
    t = torch.arange(8.)
    target = (0.02 + 0.5 * torch.exp(-((t - 5.0) / 4.0) ** 2)).reshape(-1, 1, 1).cuda()

    def step_fn(y0):
        mod.clear_tracking()
        y = ude_b200.odeint(mod, y0, t, method="rk4", options=dict(step_size=t[1] - t[0]))
        post = mod.posterior()
        return ((y[..., 1] - target) ** 2).mean() + 0.01 * (post.mean ** 2).sum() + 0.01 * torch.norm(torch.stack(mod.tracker)) / B ** 0.5
    return mod, y0s, step_fn


@pytest.mark.parametrize("grad_lim", [None, 1e9, 1e-12])
def test_graphed_iteration_equals_eager(grad_lim):
    import ude_b200
    mod_e, y0s, step_e = _setup()
    opt_e = torch.optim.Adam(mod_e.parameters(), lr=1e-3, capturable=True)
    eager = ude_b200.GraphedIteration(step_e, mod_e.parameters(), opt_e, grad_lim=grad_lim)
    mod_g, _, step_g = _setup()
    opt_g = torch.optim.Adam(mod_g.parameters(), lr=1e-3, capturable=True)
    graphed = ude_b200.GraphedIteration(step_g, mod_g.parameters(), opt_g, grad_lim=grad_lim)
    losses_e, losses_g = [], []
    for y0 in y0s:
        losses_e.append(float(eager._iteration([y0])[0]))        # the same body, eagerly
        losses_g.append(float(graphed(y0)))
    assert losses_e == pytest.approx(losses_g, rel=1e-6)
    for pe, pg in zip(mod_e.parameters(), mod_g.parameters()):
        assert torch.allclose(pe, pg, rtol=1e-6, atol=1e-9)
    if grad_lim == 1e-12:        # gated: nothing moved
        ref, _, _ = _setup()
        for p0, pg in zip(ref.parameters(), mod_g.parameters()):
            assert torch.equal(p0, pg)
        assert float(graphed.last_ok) == 0.0
    else:
        assert losses_g[-1] != losses_g[0]


def test_epoch_driver_and_recapture_for_a_new_horizon():
    """The curriculum grows the horizon (run_ode.py:147-164): a new time grid is a new graph, captured in mid-training
    without disturbing the optimizer state."""
    import ude_b200
    mod, y0s, _ = _setup(B=512)
    opt = torch.optim.Adam(mod.parameters(), lr=1e-3, capturable=True)
    horizons = {}

    def make(k):
        t = torch.arange(float(k + 1))

        def step_fn(y0):
            mod.clear_tracking()
            y = ude_b200.odeint(mod, y0, t, method="rk4", options=dict(step_size=1.0))
            return (y[..., 1] ** 2).mean()
        return step_fn

    it = None
    stats = []
    for k in (3, 5):
        fn = make(k)
        it = ude_b200.GraphedIteration(fn, mod.parameters(), opt, grad_lim=5000.0)
        stats.append(ude_b200.run_epoch(it, [(y,) for y in y0s], tag=k))
    assert all(s["iterations"] == 4 and s["steps_taken"] == 4 for s in stats)
    assert float(opt.state[next(iter(mod.parameters()))]["step"]) == 8.0


def test_full_fused_vae_style_iteration_in_one_graph():
    """Everything either side of the solve at once (SURVEY 8 f1 + f2 + f4), shaped like lib/VAE.py:117-223: an encoder stub gives
    (mean, std); reparam + solve + Decoder + latent_init_loss run fused (odeint_decoded with reparam=...); mse + nll over the
    sample axis + kl_p + Fa_norm + reg_loss form the loss; backward and a gated Adam step follow -- captured in ONE CUDA graph and
    compared with the same iteration run eagerly through the un-fused torch ops."""
    import ude_b200
    from torch.distributions import Normal
    torch.manual_seed(0)
    S, b, L, T = 16, 24, 8, 6
    t = torch.arange(float(T))

    def build():
        torch.manual_seed(1)
        ode = ude_b200.FaFp(1, L, net_sizes=[64, 64, 32], aug_net_sizes=[64, 64])
        with torch.no_grad():
            ode.aug_net[-1].weight.mul_(0.05); ode.aug_net[-1].bias.mul_(0.05)
        dec = ude_b200.Decoder(1, L, 1)
        enc = torch.nn.Linear(5, 2 * (L - 1))
        return ode.cuda(), dec.cuda(), enc.cuda()

    x = torch.randn(b, 5).cuda()
    y_true = (0.05 + 0.3 * torch.rand(b, T, 1)).cuda()
    eps = torch.randn(S, b, 1, L - 1).cuda()
    base = torch.tensor([0.9, 0.02] + [0.0] * (L - 3)).cuda()

    def encode(enc):
        h = enc(x)
        mean = base + 0.01 * h[:, :L - 1].reshape(b, 1, L - 1)
        std = 0.02 + 0.01 * torch.sigmoid(h[:, L - 1:]).reshape(b, 1, L - 1)
        return mean, std

    def loss_tail(y, reg, ode):                      # lib/VAE.py:142-196 on y_pred [b, S, T, R]
        y_pred = y.reshape(T, S, b, 1).permute(2, 1, 0, 3)
        mse = torch.mean(torch.square(y_pred - y_true.unsqueeze(1)))
        nll = -Normal(y_pred.mean(1), y_pred.std(1), validate_args=False).log_prob(y_true).mean()
        post = ode.posterior()
        kl_p = 0.01 * ((post.mean - 0.5) ** 2).sum() + 0.01 * (post.stddev ** 2).sum()
        return mse + 1e-3 * nll + kl_p + 0.1 * torch.norm(torch.stack(ode.tracker)) / (S * b) ** 0.5 + 0.1 * reg / (S * b)

    def fused_step(ode, dec, enc):
        def fn(_):
            mean, std = encode(enc)
            ode.clear_tracking()
            y, reg = ude_b200.odeint_decoded(ode, dec, None, t, options=dict(step_size=1.0), reparam=(eps, std, mean))
            return loss_tail(y, reg, ode)
        return fn

    def plain_step(ode, dec, enc):
        def fn(_):
            mean, std = encode(enc)
            z = eps * std + mean
            z = torch.cat([z[..., :2].abs(), (1 - z[..., :2].abs().sum(-1)).unsqueeze(-1), z[..., 2:]], -1).reshape(S * b, 1, L) + 1e-5
            ode.clear_tracking()
            lat = ude_b200.odeint(ode, z, t, method="rk4", options=dict(step_size=1.0))
            x3 = lat[..., :3]
            reg = (torch.where(x3 < 0, x3.abs(), torch.zeros_like(x3)) + torch.where(x3 > 1, (1 - x3).abs(), torch.zeros_like(x3))).sum()
            return loss_tail(dec(x3), reg, ode)
        return fn

    dummy = torch.zeros(1, device="cuda")
    runs = {}
    for name, make, graphed in (("fused-graph", fused_step, True), ("plain-eager", plain_step, False)):
        ode, dec, enc = build()
        params = [*ode.parameters(), *dec.parameters(), *enc.parameters()]
        opt = torch.optim.Adam(params, lr=1e-3, capturable=True)
        it = ude_b200.GraphedIteration(make(ode, dec, enc), params, opt, grad_lim=300.0)
        losses = [float(it(dummy)) if graphed else float(it._iteration([dummy])[0]) for _ in range(4)]
        runs[name] = (losses, [p.detach().clone() for p in params])
    assert runs["fused-graph"][0] == pytest.approx(runs["plain-eager"][0], rel=2e-5)
    assert runs["fused-graph"][0][-1] < runs["fused-graph"][0][0]
    for pa, pb in zip(runs["fused-graph"][1], runs["plain-eager"][1]):
        assert torch.allclose(pa, pb, rtol=1e-3, atol=2e-6)          # four Adam steps: lr-sized moves, same direction
