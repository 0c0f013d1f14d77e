"""Weight-set groups (BASELINE configs[4]: tuning/tune_node.py-style sweeps — G independent Fp models, each with its
own 64 x 32 trajectories, trained side by side).  ``odeint_groups`` must be exactly G separate ``odeint`` calls:
bit-identical to them on the GPU, and within the 1e-5 parity bar of the oracle."""
import pytest
import torch

from tests.util import RTOL_PARITY, flat_grad, module_from_spec, rel_err

pytestmark = pytest.mark.gpu


def _models(spec, G, seed):
    from oracle.ude_rhs import init_theta
    thetas = [init_theta(spec, seed=seed + g) for g in range(G)]
    return thetas, [module_from_spec(spec, th, 1.0, "cuda") for th in thetas]


@pytest.mark.parametrize("kernels", ["ffma", "tensor"])
@pytest.mark.parametrize("kind,L,nets,aug,G,Bg", [
    ("Fp", 8, (32, 32), (), 5, 300),            # tune_node.py:170 widths, ragged tiles
    ("Fp", 16, (32, 32), (), 3, 64),
    ("FaFp", 8, (64, 64, 32), (64, 64), 4, 130),
])
def test_groups_equal_separate_solves_and_oracle(kind, L, nets, aug, G, Bg, kernels):
    import ude_b200
    from oracle import odeint_ref
    from oracle.ude_rhs import OracleFunc, UdeSpec, synth_y0
    spec = UdeSpec(kind, 1, L, nets, aug)
    thetas, mods = _models(spec, G, seed=40)
    t = torch.linspace(1, 9, 9) / 7                       # off-grid outputs, short last step (tune_node.py:192,204)
    h = 0.5
    y0 = torch.stack([synth_y0(Bg, 1, L, seed=70 + g) for g in range(G)])
    Gc = torch.randn(G, t.numel(), Bg, 1, L, generator=torch.Generator().manual_seed(9))

    # one grouped call
    y0g = y0.cuda().requires_grad_(True)
    for m in mods:
        m.clear_tracking()
    y = ude_b200.odeint_groups(mods, y0g, t, method="rk4", options=dict(step_size=h, kernels=kernels))
    assert y.shape == (G, t.numel(), Bg, 1, L)
    (y * Gc.cuda()).sum().backward()
    grouped = [(y[g].detach().clone(), y0g.grad[g].clone(), flat_grad(mods[g]).clone()) for g in range(G)]
    side_lens = [(len(m.params), len(getattr(m, "tracker", []))) for m in mods]

    for g in range(G):
        # the same model through the single-set entry point: same kernels, same tiles -> same bits
        mod = module_from_spec(spec, thetas[g], 1.0, "cuda")
        y0s = y0[g].cuda().requires_grad_(True)
        ys = ude_b200.odeint(mod, y0s, t, method="rk4", options=dict(step_size=h, kernels=kernels))      # the same engine as the grouped launch
        (ys * Gc[g].cuda()).sum().backward()
        # per-trajectory results do not depend on the tiling: same bits; dtheta is a sum over tiles whose size follows the
        # TOTAL batch in the grouped launch (64 instead of 32 trajectories): same value to fp32 summation round-off
        assert torch.equal(ys, grouped[g][0])
        assert torch.equal(y0s.grad, grouped[g][1])
        assert rel_err(grouped[g][2], flat_grad(mod)) <= 2e-6
        # and the oracle
        f = OracleFunc(spec, thetas[g])
        y0r = y0[g].clone().requires_grad_(True)
        yr = odeint_ref.odeint(f, y0r, t, method="rk4", options=dict(step_size=h))
        (yr * Gc[g]).sum().backward()
        assert rel_err(grouped[g][0], yr.detach()) <= RTOL_PARITY
        assert rel_err(grouped[g][1], y0r.grad) <= RTOL_PARITY
        assert rel_err(grouped[g][2], f.theta.grad) <= RTOL_PARITY
    assert all(a == (1 if spec.has_rate else 0) for a, _ in side_lens)


def test_group_side_channels_are_per_model():
    import ude_b200
    from oracle.ude_rhs import UdeSpec, synth_y0
    spec = UdeSpec("Fp", 1, 8, (32, 32), ())
    thetas, mods = _models(spec, 3, seed=5)
    t = torch.arange(4.)
    y0 = torch.stack([synth_y0(200, 1, 8, seed=g) for g in range(3)]).cuda()
    for m in mods:
        m.clear_tracking()
    with torch.no_grad():
        ude_b200.odeint_groups(mods, y0, t, method="rk4", options=dict(step_size=1.0))
    posts = [m.posterior() for m in mods]
    for g in range(3):
        mod = module_from_spec(spec, thetas[g], 1.0, "cuda")
        mod.clear_tracking()
        with torch.no_grad():
            ude_b200.odeint(mod, y0[g], t, method="rk4", options=dict(step_size=1.0, kernels="ffma"))
        p = mod.posterior()
        assert rel_err(posts[g].mean, p.mean) <= 1e-6 and rel_err(posts[g].stddev, p.stddev) <= 1e-5
    assert not torch.equal(posts[0].mean, posts[1].mean)


@pytest.mark.parametrize("kernels,path", [("ffma", 2), (None, 3)])
def test_many_groups_fan_out(kernels, path):
    """64 groups x 2048 trajectories (the per-GPU share of configs[4]) in ONE forward and ONE backward launch (plan path 2: FFMA
    engine, 3: tcgen05 engine = the default): finite, and group 17 equals its own single-set solve."""
    import ctypes as C

    import ude_b200
    from ude_b200 import _lib
    from oracle.ude_rhs import UdeSpec, synth_y0
    spec = UdeSpec("Fp", 1, 8, (32, 32), ())
    G, Bg = 64, 2048
    thetas, mods = _models(spec, G, seed=100)
    plan = _lib.UdePlan()
    info = ude_b200.describe_func(mods[0])
    info.kernels = kernels
    desc = info.desc(G * Bg, n_groups=G)
    assert ude_b200.lib().ude_plan(C.byref(desc), 0, 8, C.byref(plan)) == 0 and plan.path == path, "not the single-launch grouped path"
    t = torch.linspace(1, 56, 56) / 7
    y0 = synth_y0(G * Bg, 1, 8, seed=3).reshape(G, Bg, 1, 8).cuda().requires_grad_(True)
    y = ude_b200.odeint_groups(mods, y0, t, method="rk4", options=dict(step_size=1.0, kernels=kernels))
    (y[..., :3] ** 2).sum().backward()
    assert torch.isfinite(y).all() and torch.isfinite(y0.grad).all()
    mod = module_from_spec(spec, thetas[17], 1.0, "cuda")
    y0s = y0[17].detach().clone().requires_grad_(True)
    ys = ude_b200.odeint(mod, y0s, t, method="rk4", options=dict(step_size=1.0, kernels=kernels))
    (ys[..., :3] ** 2).sum().backward()
    assert torch.equal(ys, y[17])
    assert torch.equal(y0s.grad, y0.grad[17])
    assert rel_err(flat_grad(mods[17]), flat_grad(mod)) <= 2e-6


def test_group_engines_agree_on_every_group_and_flat_theta():
    """1024 tiles over 148 CTAs (a CTA changes group at every tile): the tcgen05 and the FFMA grouped launches must agree on EVERY
    group's trajectories and gradients, and the one-leaf parameter form (flatten_group_parameters) must give the same gradients."""
    import ude_b200
    from oracle.ude_rhs import UdeSpec, synth_y0
    spec = UdeSpec("Fp", 1, 8, (32, 32), ())
    G, Bg = 64, 2048
    t = torch.arange(5.)
    y0 = synth_y0(G * Bg, 1, 8, seed=8).reshape(G, Bg, 1, 8).cuda()
    Gc = torch.randn(G, 5, Bg, 1, 8, generator=torch.Generator().manual_seed(4)).cuda()
    res = {}
    for kernels in ("tensor", "ffma"):
        thetas, mods = _models(spec, G, seed=300)
        y0g = y0.clone().requires_grad_(True)
        y = ude_b200.odeint_groups(mods, y0g, t, method="rk4", options=dict(step_size=1.0, kernels=kernels))
        (y * Gc).sum().backward()
        res[kernels] = (y.detach(), y0g.grad, torch.stack([flat_grad(m) for m in mods]), torch.stack([m.posterior().mean for m in mods]))
    bad_rows = 0
    for g in range(G):
        assert rel_err(res["tensor"][0][g], res["ffma"][0][g]) <= RTOL_PARITY, g
        assert rel_err(res["tensor"][3][g], res["ffma"][3][g]) <= RTOL_PARITY, g
        # dL/dy0 row by row: the RHS has kinks (beta = |o|, lib/models.py:135) -- a trajectory whose head output passes within
        # an ulp of zero takes the other branch in one of the two families (see test_tensor_and_ffma_families_agree_at_scale)
        a, b = res["tensor"][1][g], res["ffma"][1][g]
        row_err = (a - b).abs().amax(dim=(1, 2)) / float(b.abs().max())
        kinked = row_err > RTOL_PARITY
        bad_rows += int(kinked.sum())
        if not bool(kinked.any()):      # dL/dtheta sums over the trajectories: comparable when no row of the group sits on a kink
            assert rel_err(res["tensor"][2][g], res["ffma"][2][g]) <= RTOL_PARITY, g
    assert bad_rows <= max(1, G * Bg // 20_000), bad_rows
    thetas, mods = _models(spec, G, seed=300)
    flat = ude_b200.flatten_group_parameters(mods)
    assert flat.shape == (G, thetas[0].numel()) and torch.equal(flat[5].detach().cpu(), thetas[5])
    y0g = y0.clone().requires_grad_(True)
    y = ude_b200.odeint_groups(mods, y0g, t, method="rk4", options=dict(step_size=1.0), theta=flat)
    (y * Gc).sum().backward()
    assert torch.equal(y.detach(), res["tensor"][0]) and torch.equal(flat.grad, res["tensor"][2])
    with torch.no_grad():
        flat.mul_(0.5)                      # the modules see the buffer: their parameters are views of its rows
    assert torch.equal(mods[3].Fp_net[1].weight.reshape(-1), flat[3, :mods[3].Fp_net[1].weight.numel()].detach())


def test_group_architecture_mismatch_is_an_error():
    import ude_b200
    from oracle.ude_rhs import UdeSpec, init_theta
    a = UdeSpec("Fp", 1, 8, (32, 32), ())
    b = UdeSpec("Fp", 1, 8, (20, 20), ())
    mods = [module_from_spec(a, init_theta(a, 0), 1.0, "cuda"), module_from_spec(b, init_theta(b, 0), 1.0, "cuda")]
    with pytest.raises(ValueError):
        ude_b200.odeint_groups(mods, torch.zeros(2, 4, 1, 8, device="cuda"), torch.arange(3.), options=dict(step_size=1.0))
