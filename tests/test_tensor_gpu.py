"""GPU parity of the tcgen05 (tensor-core) RK4 kernels — the family the library picks for the
BASELINE configs[3] ensemble (B >= 148 x 128 trajectories, R = 1, nets <= 64 wide).  Same
checks as test_rk4_gpu.py (committed goldens generated with the REAL reference modules, oracle
on fresh seeded inputs, ragged tiles, determinism), with the kernel family pinned through
``options['kernels']`` so small batches reach the tensor path too.  Tolerance: BASELINE
north_star, max|x - ref| / max|ref| <= 1e-5 per tensor."""
import ctypes as C

import pytest
import torch

from tests.util import RTOL_PARITY, flat_grad, golden_names, load_golden, module_from_spec, rel_err

pytestmark = pytest.mark.gpu

SM_TILES = 148 * 128        # one 128-trajectory tile per SM


def _path(mod, B, kernels=None):
    """Kernel family the library would run: 0 = FFMA, 1 = tensor, < 0 = unsupported."""
    import ude_b200
    info = ude_b200.describe_func(mod)
    info.kernels = kernels
    return int(ude_b200.lib().ude_rk4_path(C.byref(info.desc(B))))


def _tensor_goldens():
    out = []
    for name in golden_names("rk4_"):
        g = load_golden(name)
        s = g["spec"]
        nets = [n for n in (s.net_sizes, s.aug_sizes) if len(n)]
        if s.n_regions == 1 and s.latent_dim in (4, 8, 16) and all(max(n) <= 64 for n in nets):
            out.append(name)
    return out


def _solve(g, kernels="tensor"):
    import ude_b200
    mod = module_from_spec(g["spec"], torch.from_numpy(g["theta"]), g["fa_w"], "cuda")
    y0 = torch.from_numpy(g["y0"]).cuda().requires_grad_(True)
    assert _path(mod, y0.shape[0], kernels) == 1, "golden case is not covered by the tensor path"
    mod.clear_tracking()
    y = ude_b200.odeint(mod, y0, torch.from_numpy(g["t"]), method="rk4",
                        options=dict(step_size=float(g["step_size"]), kernels=kernels))
    return mod, y0, y


def test_some_goldens_reach_the_tensor_path():
    assert len(_tensor_goldens()) >= 3


@pytest.mark.parametrize("name", _tensor_goldens())
def test_golden_forward_and_side_channels(name):
    g = load_golden(name)
    mod, y0, y = _solve(g)
    assert rel_err(y, torch.from_numpy(g["y"])) <= RTOL_PARITY
    spec, side_ref = g["spec"], torch.from_numpy(g["side"])
    n = side_ref[5]
    if spec.has_rate:
        post = mod.posterior()
        mean_ref = side_ref[0:2] / n
        std_ref = torch.sqrt((side_ref[2:4] - n * mean_ref ** 2) / (n - 1))
        assert rel_err(post.mean, mean_ref) <= RTOL_PARITY
        assert rel_err(post.stddev, std_ref) <= 1e-4     # std is a difference of sums: looser
    if spec.has_aug:
        assert rel_err(torch.norm(torch.stack(mod.tracker)), torch.sqrt(side_ref[4])) <= RTOL_PARITY


@pytest.mark.parametrize("name", _tensor_goldens())
def test_golden_backward(name):
    g = load_golden(name)
    mod, y0, y = _solve(g)
    G = torch.from_numpy(g["G"]).cuda()
    c = torch.from_numpy(g["c_side"]).cuda()
    loss = (y.double() * G.double()).sum()              # same loss as make_golden.py
    spec, n = g["spec"], float(g["side"][5])
    if spec.has_rate:
        rows = torch.stack(mod.params).reshape(-1, 2).double()
        mean, var = rows.mean(0), rows.var(0)
        loss = loss + (c[0:2] * mean * n).sum() + (c[2:4] * (var * (n - 1) + n * mean * mean)).sum()
    if spec.has_aug:
        loss = loss + c[4] * torch.stack(mod.tracker).double().pow(2).sum()
    loss.backward()
    e_y0 = rel_err(y0.grad, torch.from_numpy(g["dy0"]))
    e_th = rel_err(flat_grad(mod), torch.from_numpy(g["dtheta"]))
    print(f"{name}: dy0 {e_y0:.3e} dtheta {e_th:.3e}")
    assert e_y0 <= RTOL_PARITY and e_th <= RTOL_PARITY


def _oracle_case(spec, B, t, h, seed):
    from oracle import odeint_ref
    from oracle.ude_rhs import OracleFunc, init_theta, synth_y0
    theta = init_theta(spec, seed=seed)
    y0 = synth_y0(B, spec.n_regions, spec.latent_dim, seed=seed + 1)
    G = torch.randn(t.numel(), B, spec.n_regions, spec.latent_dim, generator=torch.Generator().manual_seed(seed + 2))
    f = OracleFunc(spec, theta)
    y0r = y0.clone().requires_grad_(True)
    yr = odeint_ref.odeint(f, y0r, t, method="rk4", options=dict(step_size=h))
    (yr * G).sum().backward()
    return theta, y0, G, yr.detach(), y0r.grad, f.theta.grad


def _fused_case(spec, theta, y0, G, t, h, kernels):
    import ude_b200
    mod = module_from_spec(spec, theta, 1.0, "cuda")
    y0g = y0.cuda().requires_grad_(True)
    y = ude_b200.odeint(mod, y0g, t, method="rk4", options=dict(step_size=h, kernels=kernels))
    (y * G.cuda()).sum().backward()
    return y.detach(), y0g.grad, flat_grad(mod)


@pytest.mark.parametrize("B", [1, 127, 128, 129, 300, SM_TILES + 77])
def test_ragged_tiles_match_oracle(B):
    """Partial tiles (B not a multiple of 128), a single trajectory, and one tile more than the SM count."""
    from oracle.ude_rhs import UdeSpec
    spec = UdeSpec("FaFp", 1, 8, (64, 64, 32), (64, 64))
    t = torch.arange(3.)
    theta, y0, G, yr, dy0r, dthr = _oracle_case(spec, B, t, 1.0, seed=B)
    y, dy0, dth = _fused_case(spec, theta, y0, G, t, 1.0, "tensor")
    assert rel_err(y, yr) <= RTOL_PARITY
    assert rel_err(dy0, dy0r) <= RTOL_PARITY
    assert rel_err(dth, dthr) <= RTOL_PARITY


@pytest.mark.parametrize("kind,L,nets,aug", [
    ("Fp", 8, (64, 64), ()), ("Fp", 4, (32, 32), ()), ("Fa", 16, (), (64, 64)), ("FaFp", 4, (32, 32), (32, 32)),
    ("FaFp", 16, (32, 32), (32, 32)), ("Fp", 16, (64, 64, 32), ()), ("Fp", 8, (20, 20), ()),
])
def test_model_shapes_match_oracle(kind, L, nets, aug):
    """Single-net kinds (only one chain runs), tune_node-style widths (tuning/tune_node.py:170), every latent_dim
    the tensor path instantiates, widths that need zero padding, and an off-grid output grid (linear
    interpolation + a short last step, tuning/tune_node.py:192,204)."""
    from oracle.ude_rhs import UdeSpec
    spec = UdeSpec(kind, 1, L, nets, aug)
    t = torch.linspace(1, 9, 9) / 7
    theta, y0, G, yr, dy0r, dthr = _oracle_case(spec, 333, t, 0.5, seed=11)
    y, dy0, dth = _fused_case(spec, theta, y0, G, t, 0.5, "tensor")
    assert rel_err(y, yr) <= RTOL_PARITY
    assert rel_err(dy0, dy0r) <= RTOL_PARITY
    assert rel_err(dth, dthr) <= RTOL_PARITY


def test_largest_model_falls_back_to_the_ffma_engine():
    """FaFp with latent_dim 16 AND the 64-wide nets needs 2.7 KB more shared memory than an SM has (the fp32 first-layer
    weights of the FMA-pipe W0^T delta0 double with D0): the library must pick the FFMA engine for it, and pinning the
    tensor family must fail loudly instead of silently running something else."""
    import ude_b200
    from oracle.ude_rhs import UdeSpec, init_theta
    spec = UdeSpec("FaFp", 1, 16, (64, 64, 32), (64, 64))
    mod = module_from_spec(spec, init_theta(spec, seed=0), 1.0, "cuda")
    assert _path(mod, 1 << 20) == 0
    assert _path(mod, 1 << 20, "tensor") < 0
    with pytest.raises(ude_b200.UdeError):
        ude_b200.odeint(mod, torch.zeros(256, 1, 16, device="cuda"), torch.arange(3.), method="rk4",
                        options=dict(step_size=1.0, kernels="tensor"))


def test_library_picks_tensor_path_for_the_ensemble_and_matches_oracle():
    """The default dispatch (no pin) is the tensor family for every covered model at every batch size (measured faster from
    B = 1 on: profiles/r02_engine_crossover.txt), and parity must hold there."""
    from oracle.ude_rhs import UdeSpec, init_theta
    spec = UdeSpec("FaFp", 1, 8, (64, 64, 32), (64, 64))
    mod = module_from_spec(spec, init_theta(spec, seed=0), 1.0, "cuda")
    assert _path(mod, SM_TILES) == 1 and _path(mod, 1) == 1 and _path(mod, 1, "ffma") == 0
    t = torch.arange(5.)
    theta, y0, G, yr, dy0r, dthr = _oracle_case(spec, SM_TILES, t, 1.0, seed=21)
    y, dy0, dth = _fused_case(spec, theta, y0, G, t, 1.0, None)
    assert rel_err(y, yr) <= RTOL_PARITY
    assert rel_err(dy0, dy0r) <= RTOL_PARITY
    assert rel_err(dth, dthr) <= RTOL_PARITY


def test_tensor_and_ffma_families_agree_at_scale():
    """No oracle at this size: the two independent kernel families must agree (both <= 1e-5 from the truth)."""
    from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
    spec = UdeSpec("FaFp", 1, 8, (64, 64, 32), (64, 64))
    theta = init_theta(spec, seed=3)
    B = 100_000
    y0 = synth_y0(B, 1, 8, seed=4)
    t = torch.arange(9.)
    G = torch.randn(9, B, 1, 8, generator=torch.Generator().manual_seed(5))
    a = _fused_case(spec, theta, y0, G, t, 1.0, "tensor")
    b = _fused_case(spec, theta, y0, G, t, 1.0, "ffma")
    assert rel_err(a[0], b[0]) <= RTOL_PARITY           # trajectories
    assert rel_err(a[2], b[2]) <= RTOL_PARITY           # dL/dtheta (a sum over trajectories)
    # dL/dy0 row by row.  The RHS has kinks (beta = |o|, gamma = |o'|, lib/models.py:135; the [-1, 2] clamp, :130-146):
    # a trajectory whose head output passes within one ulp of zero takes the other branch in one of the two
    # families and its own gradient row differs at O(1); every other row must agree to the parity tolerance.
    scale = float(b[1].abs().max())
    row_err = (a[1] - b[1]).abs().amax(dim=(1, 2)) / scale
    bad = int((row_err > RTOL_PARITY).sum())
    print(f"rows over tolerance: {bad} of {B}; worst {float(row_err.max()):.3e}")
    assert bad <= max(1, B // 20_000)


def test_tensor_backward_is_deterministic():
    """Gradient slabs are per CTA, every slab address has ONE writer, slabs are summed in a fixed order."""
    from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
    spec = UdeSpec("FaFp", 1, 8, (64, 64, 32), (64, 64))
    theta = init_theta(spec, seed=7)
    y0 = synth_y0(40_000, 1, 8, seed=3)
    t = torch.arange(5.)
    G = torch.randn(5, 40_000, 1, 8, generator=torch.Generator().manual_seed(6))
    r1 = _fused_case(spec, theta, y0, G, t, 1.0, "tensor")
    r2 = _fused_case(spec, theta, y0, G, t, 1.0, "tensor")
    for u, v in zip(r1, r2):
        assert torch.equal(u, v)


def test_clamped_trajectories():
    """States outside [-1, 2] zero the RHS component-wise (lib/models.py:130-146) on the tensor path too."""
    from oracle.ude_rhs import UdeSpec
    spec = UdeSpec("FaFp", 1, 8, (64, 64, 32), (64, 64))
    t = torch.arange(4.)
    theta, y0, G, _, _, _ = _oracle_case(spec, 256, t, 1.0, seed=31)
    y0 = y0.clone()
    y0[::3, 0, 0] = 2.5
    y0[1::5, 0, 1] = -1.5
    from oracle import odeint_ref
    from oracle.ude_rhs import OracleFunc
    f = OracleFunc(spec, theta)
    y0r = y0.clone().requires_grad_(True)
    yr = odeint_ref.odeint(f, y0r, t, method="rk4", options=dict(step_size=1.0))
    (yr * G).sum().backward()
    y, dy0, dth = _fused_case(spec, theta, y0, G, t, 1.0, "tensor")
    assert rel_err(y, yr.detach()) <= RTOL_PARITY
    assert rel_err(dy0, y0r.grad) <= RTOL_PARITY
    assert rel_err(dth, f.theta.grad) <= RTOL_PARITY


@pytest.mark.parametrize("kind,L,nets,aug", [
    ("FaFp", 8, (16, 16), (16, 16)),            # thinnest legal nets: every phase is short, hand-offs are back to back
    ("FaFp", 4, (32, 16), (64, 32)),            # unequal chains: the lock step runs out of rate layers first
    ("FaFp", 16, (32, 32), (32, 32)),
    ("Fp", 8, (64, 16, 64), ()),
    ("Fa", 8, (), (32, 64, 32)),
    ("FaFp", 8, (64, 64, 32), (64, 64)),
])
def test_shapes_agree_with_ffma_family_over_several_tiles_per_cta(kind, L, nets, aug):
    """Timing-sensitive coverage: 40 000 trajectories = 313 tiles on 148 CTAs (2-3 tiles each, ragged tail), short
    nets, both kernel families on the same inputs, and a bit-identical rerun of the tensor family."""
    from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
    spec = UdeSpec(kind, 1, L, nets, aug)
    theta = init_theta(spec, seed=13)
    B = 40_000
    y0 = synth_y0(B, 1, L, seed=14)
    t = torch.arange(5.)
    G = torch.randn(5, B, 1, L, generator=torch.Generator().manual_seed(15))
    a = _fused_case(spec, theta, y0, G, t, 1.0, "tensor")
    a2 = _fused_case(spec, theta, y0, G, t, 1.0, "tensor")
    b = _fused_case(spec, theta, y0, G, t, 1.0, "ffma")
    for u, v in zip(a, a2):
        assert torch.equal(u, v)
    assert rel_err(a[0], b[0]) <= RTOL_PARITY
    assert rel_err(a[2], b[2]) <= RTOL_PARITY
    scale = float(b[1].abs().max())
    row_err = (a[1] - b[1]).abs().amax(dim=(1, 2)) / scale
    assert int((row_err > RTOL_PARITY).sum()) <= max(1, B // 20_000)
