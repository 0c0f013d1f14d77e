"""CPU tests of the oracle itself (no GPU): against the committed golden fixtures (generated with
the REAL reference modules), against closed forms, and self-consistency of the Dopri5 restatement."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import odeint_ref
from oracle.ude_rhs import OracleFunc, UdeSpec, init_theta, synth_y0, ude_rhs
from tests.util import golden_names, load_golden, rel_err


class _Decay(torch.nn.Module):
    def forward(self, t, y):
        return -y


@pytest.mark.parametrize("name", golden_names("rk4_"))
def test_oracle_reproduces_rk4_golden(name):
    """Restated RHS + restated solver == real lib.models classes + restated solver (bit for bit
    forward; gradients to fp32 round-off since the flat-theta graph differs)."""
    g = load_golden(name)
    f = OracleFunc(g["spec"], torch.from_numpy(g["theta"]), g["fa_w"])
    y0 = torch.from_numpy(g["y0"]).requires_grad_(True)
    t = torch.from_numpy(g["t"])
    y = odeint_ref.odeint(f, y0, t, method="rk4", options=dict(step_size=torch.tensor(float(g["step_size"]))))
    assert torch.equal(y.detach(), torch.from_numpy(g["y"]))
    side = f.side_sums()
    assert rel_err(side, torch.from_numpy(g["side"])) < 1e-12
    loss = (y.double() * torch.from_numpy(g["G"]).double()).sum() + (side[:5] * torch.from_numpy(g["c_side"])).sum()
    loss.backward()
    assert rel_err(y0.grad, torch.from_numpy(g["dy0"])) < 1e-6
    assert rel_err(f.theta.grad, torch.from_numpy(g["dtheta"])) < 1e-6


@pytest.mark.parametrize("name", golden_names("dopri5"))
def test_oracle_reproduces_dopri5_golden(name):
    g = load_golden(name)
    f = OracleFunc(g["spec"], torch.from_numpy(g["theta"]), g["fa_w"])
    stats = odeint_ref.Dopri5Stats()
    opts = dict(per_trajectory=True, stats=stats)
    if float(g["first_step"]) > 0:
        opts["first_step"] = float(g["first_step"])
    y = odeint_ref.odeint(f, torch.from_numpy(g["y0"]), torch.from_numpy(g["t"]), rtol=float(g["rtol"]),
                          atol=float(g["atol"]), options=opts)
    assert torch.equal(y.detach(), torch.from_numpy(g["y"]))
    assert np.array_equal(stats.n_accept.numpy(), g["n_accept"])
    assert np.array_equal(stats.n_reject.numpy(), g["n_reject"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/lib"), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("kind,R,L,kw", [
    ("Fp", 1, 8, dict(net_sizes=[64, 64, 32])), ("Fa", 3, 8, dict(aug_net_sizes=[64, 64])),
    ("FaFp", 2, 8, dict(net_sizes=[64, 64, 32], aug_net_sizes=[64, 64])), ("FaFp", 1, 4, {}), ("Fp", 1, 16, {}),
])
def test_rhs_restatement_matches_real_reference_modules(kind, R, L, kw):
    """oracle.ude_rhs vs /root/reference/lib/models.py, bit for bit incl. side channels + clamp."""
    import sys
    sys.path.insert(0, "/root/reference")
    import lib.models as ref_models
    from oracle.ude_rhs import theta_from_module
    torch.manual_seed(0)
    mod = getattr(ref_models, kind)(R, L, **kw)
    if kind == "FaFp":
        mod.Fa_w = 0.37
    spec, theta, fa_w = theta_from_module(mod)
    x = synth_y0(9, R, L, seed=4)
    x[0, 0, 0] = 2.5
    x[1, 0, 1] = -1.5
    x[2, R - 1, 2] = 3.0
    f, p, fa = ude_rhs(spec, theta, x, fa_w)
    ref = mod(torch.tensor(0.0), x)
    assert torch.equal(f, ref)
    if p is not None:
        assert torch.equal(p, mod.params[0])
    if fa is not None:
        assert torch.equal(fa, mod.tracker[0])


def test_rk4_three_eighths_rule_order_and_closed_form():
    y0 = torch.tensor([[1.0, 2.0]], dtype=torch.float64)
    t = torch.linspace(0, 2, 5, dtype=torch.float64)
    errs = []
    for h in (0.1, 0.05):
        y = odeint_ref.odeint(_Decay(), y0, t, method="rk4", options=dict(step_size=h))
        errs.append(float((y[-1] - y0 * math.exp(-2)).abs().max()))
    assert errs[0] < 1e-6
    assert 12 < errs[0] / errs[1] < 20        # 4th order: halving h divides the error by ~16
    # one step equals the 3/8-rule polynomial 1 - h + h^2/2 - h^3/6 + h^4/24 for y' = -y
    h = 0.3
    y = odeint_ref.odeint(_Decay(), y0, torch.tensor([0.0, h], dtype=torch.float64), method="rk4")
    assert torch.allclose(y[1], y0 * (1 - h + h ** 2 / 2 - h ** 3 / 6 + h ** 4 / 24), rtol=1e-14)


def test_rk4_grid_and_interpolation_rules():
    """torchdiffeq grid: ceil(span/h + 1) points, last overwritten; off-grid outputs lerp."""
    t = torch.linspace(1, 56, 56) / 7
    g = odeint_ref.rk4_grid(t, 1.0)
    assert g.numel() == 9 and float(g[-1]) == float(t[-1]) and abs(float(g[1] - g[0]) - 1.0) < 1e-6
    y0 = torch.tensor([[1.0]], dtype=torch.float32)
    y = odeint_ref.odeint(_Decay(), y0, t, method="rk4", options=dict(step_size=1.0))
    assert y.shape == (56, 1, 1)
    # between grid points the output is the chord, not the exact solution
    y_grid = odeint_ref.odeint(_Decay(), y0, g, method="rk4")
    s = (t[3] - g[0]) / (g[1] - g[0])
    assert torch.allclose(y[3], y_grid[0] + s * (y_grid[1] - y_grid[0]), atol=1e-7)


def test_dopri5_tableau_conditions():
    for a, row in zip(odeint_ref.DOPRI5_ALPHA, odeint_ref.DOPRI5_BETA):
        assert abs(sum(row) - a) < 1e-14
    assert abs(sum(odeint_ref.DOPRI5_C_SOL) - 1) < 1e-14
    assert abs(sum(odeint_ref.DOPRI5_C_ERR)) < 1e-14
    assert abs(sum(odeint_ref.DOPRI5_C_MID) - 0.5) < 1e-14
    alpha = [0.0, *odeint_ref.DOPRI5_ALPHA]
    assert abs(sum(c * a for c, a in zip(odeint_ref.DOPRI5_C_MID, alpha)) - 0.125) < 1e-14


def test_dopri5_closed_form_and_dense_output():
    y0 = torch.tensor([[1.0, 2.0]], dtype=torch.float64)
    t = torch.linspace(0, 2, 9, dtype=torch.float64)
    st = odeint_ref.Dopri5Stats()
    y = odeint_ref.odeint(_Decay(), y0, t, rtol=1e-8, atol=1e-10, options=dict(stats=st))
    exact = y0 * torch.exp(-t).reshape(-1, 1, 1)
    assert float((y - exact).abs().max()) < 1e-7
    assert st.n_accept > 5 and st.n_reject >= 0


def test_dopri5_per_trajectory_equals_batch_of_one_runs():
    """The fused kernels' semantics: per-trajectory control == torchdiffeq on each row alone."""
    spec = UdeSpec("FaFp", 1, 8, (16, 16), (16,))
    theta = init_theta(spec, seed=3).double()
    y0 = synth_y0(5, 1, 8, seed=7).double()
    t = torch.arange(0., 6.).double()
    f = OracleFunc(spec, theta)
    st = odeint_ref.Dopri5Stats()
    ya = odeint_ref.odeint(f, y0, t, rtol=1e-6, atol=1e-9, options=dict(per_trajectory=True, stats=st))
    total_side = f.side_sums()
    acc = torch.zeros_like(total_side)
    for b in range(5):
        f.clear_tracking()
        s1 = odeint_ref.Dopri5Stats()
        yb = odeint_ref.odeint(f, y0[b:b + 1], t, rtol=1e-6, atol=1e-9, options=dict(stats=s1))       # torchdiffeq's own (shared) mode
        n_calls = f.side_sums()[5]
        assert float((ya[:, b:b + 1] - yb).abs().max()) < 1e-9
        assert int(st.n_accept[b]) == s1.n_accept and int(st.n_reject[b]) == s1.n_reject
        assert int(n_calls) == 2 + 6 * (s1.n_accept + s1.n_reject)     # its lists hold EVERY evaluation (probe, rejected attempts)
        f.clear_tracking()
        odeint_ref.odeint(f, y0[b:b + 1], t, rtol=1e-6, atol=1e-9, options=dict(per_trajectory=True))   # accepted stages only
        acc += f.side_sums()
    assert rel_err(total_side, acc) < 1e-9


def test_oracle_adjoint_gradcheck_fp64():
    spec = UdeSpec("FaFp", 2, 4, (6, 5), (7,))
    theta = init_theta(spec, seed=1).double()
    y0 = synth_y0(3, 2, 4, seed=2).double().requires_grad_(True)
    t = torch.tensor([0.0, 0.4, 1.0], dtype=torch.float64)

    def fn(y0_, th_):
        def f(t_, x):
            return ude_rhs(spec, th_, x, 0.7)[0]
        return odeint_ref.odeint(f, y0_, t, method="rk4", options=dict(step_size=0.5))

    assert torch.autograd.gradcheck(fn, (y0, theta.clone().requires_grad_(True)), eps=1e-6, atol=1e-6)
