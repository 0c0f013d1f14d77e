"""Cross-check of the oracle's solver restatement against the REAL torchdiffeq.

The reference imports the third-party package ``torchdiffeq`` (un-vendored, un-pinned; PyPI 0.2.3 at the
reference's timestamps) at /root/reference/lib/VAE.py:5,137.  It is not installed in the build image and cannot be
installed offline, so ``oracle/odeint_ref.py`` restates its algorithm ("parity unpinned").  Wherever torchdiffeq IS
importable these tests pin the restatement: bit-for-bit on the three fixed-grid cases the reference uses, and to
1e-6 with identical accept / reject counts for Dopri5.  They skip (visibly) when the package is absent.
"""
import pytest
import torch

torchdiffeq = pytest.importorskip("torchdiffeq", reason="torchdiffeq is not installed (offline image): solver parity stays unpinned")

from oracle import odeint_ref  # noqa: E402
from oracle.ude_rhs import OracleFunc, UdeSpec, init_theta, synth_y0  # noqa: E402


def _func(kind="FaFp", R=1, L=8, seed=0):
    spec = UdeSpec(kind, R, L, (64, 64, 32) if kind != "Fa" else (), (64, 64) if kind != "Fp" else ())
    return spec, init_theta(spec, seed=seed)


GRIDS = {
    # lib/VAE.py:126,137 with run_ode.py:93,147: weekly outputs, h = t[1]-t[0] = 1 (a 0-dim tensor)
    "weekly_on_grid": lambda: (torch.arange(9.), lambda t: t[1] - t[0]),
    # run_ode.py:130: t = arange(N)/7, h = fl32(1/7): half the outputs sit one ulp below their grid point
    "daily_ulp": lambda: (torch.arange(57.) / 7, lambda t: t[1] - t[0]),
    # tuning/tune_node.py:192,204: 56 off-grid outputs interpolated between 9 grid states, short last step
    "offgrid_interp": lambda: (torch.linspace(1, 56, 56) / 7, lambda t: 1.0),
}


@pytest.mark.parametrize("grid", sorted(GRIDS))
def test_rk4_fixed_grid_bit_for_bit(grid):
    t, hf = GRIDS[grid]()
    spec, theta = _func()
    y0 = synth_y0(5, 1, 8, seed=3)
    outs = []
    for impl in (torchdiffeq.odeint, odeint_ref.odeint):
        f = OracleFunc(spec, theta)
        y0g = y0.clone().requires_grad_(True)
        y = impl(f, y0g, t, method="rk4", options=dict(step_size=hf(t)))
        (y * torch.linspace(-1, 1, y.numel()).reshape(y.shape)).sum().backward()
        outs.append((y.detach(), y0g.grad.clone(), f.theta.grad.clone(), len(f.params)))
    assert torch.equal(outs[0][0], outs[1][0]), "trajectories differ from torchdiffeq"
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2]), "gradients differ"
    assert outs[0][3] == outs[1][3], "number of RHS evaluations differs"


@pytest.mark.parametrize("B", [1, 6])
def test_dopri5_shared_step_matches(B):
    """torchdiffeq's own batch semantics (one step size for the whole tensor) vs the oracle's default mode."""
    spec, theta = _func(seed=8)
    y0 = synth_y0(B, 1, 8, seed=18)
    t = torch.arange(9.)
    f_ref, f_or = OracleFunc(spec, theta), OracleFunc(spec, theta)
    y_ref = torchdiffeq.odeint(f_ref, y0, t, rtol=1e-5, atol=1e-9)
    st = odeint_ref.Dopri5Stats()
    y_or = odeint_ref.odeint(f_or, y0, t, rtol=1e-5, atol=1e-9, options=dict(stats=st))
    assert float((y_ref - y_or).abs().max() / y_ref.abs().max()) <= 1e-6
    # accepted + rejected attempts: every attempt costs 6 evaluations, + f0 + the initial-step probe
    n_attempts = st.n_accept + st.n_reject
    assert f_ref.n_evals == 2 + 6 * n_attempts, (f_ref.n_evals, st.n_accept, st.n_reject)


def test_version_recorded():
    print("torchdiffeq", torchdiffeq.__version__)
