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
