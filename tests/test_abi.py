"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/ude_b200.h
declares, and its host-only entry points (grid construction, sizes, planning, error codes) agree
with the oracle.  No compute call needs a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from oracle import odeint_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "ude_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ude_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(built_lib, name), f"{name} declared in the header but not exported"
    import ude_b200._lib as L
    assert declared == set(L.EXPORTS)
    assert built_lib.ude_abi_version() == 1
    assert built_lib.ude_strerror(0) == b"ok" and b"workspace" in built_lib.ude_strerror(-3)


GRIDS = [
    (torch.arange(9.), 1.0), (torch.arange(57.) / 7, None), (torch.arange(37.) / 7, None),
    (torch.arange(44.) / 7, None), (torch.arange(86.) / 7, None), (torch.linspace(1, 56, 56) / 7, 1.0),
    (torch.arange(4.), 0.5), (torch.tensor([0.0, 0.3, 0.9, 2.0]), 0.7), (torch.tensor([0.0, 1e-3, 5.0]), 2.0),
]


@pytest.mark.parametrize("case", range(len(GRIDS)))
def test_host_grid_is_bit_exact_vs_oracle(built_lib, case):
    """ude_rk4_grid == torchdiffeq's fp32 grid constructor + output rule (as restated by the oracle)."""
    import ude_b200
    t, h = GRIDS[case]
    hh = (t[1] - t[0]) if h is None else h
    dt, out_step, out_kind, out_w = ude_b200.rk4_grid_tables(t.numpy(), float(hh))
    grid = odeint_ref.rk4_grid(t, hh)
    assert np.array_equal(dt, (grid[1:] - grid[:-1]).numpy())
    # replay the oracle's output loop
    j = 1
    for n, (t0, t1) in enumerate(zip(grid[:-1], grid[1:])):
        while j < len(t) and t1 >= t[j]:
            assert out_step[j] == n
            if t[j] == t0:
                assert out_kind[j] == 0
            elif t[j] == t1:
                assert out_kind[j] == 1
            else:
                assert out_kind[j] == 2
                assert np.float32(out_w[j]) == ((t[j] - t0) / (t1 - t0)).numpy()
            j += 1
    assert j == len(t) and out_step[0] == -1


def test_grid_without_step_size_is_t_itself(built_lib):
    import ude_b200
    t = torch.tensor([0.0, 0.5, 1.5, 4.0])
    dt, out_step, out_kind, _ = ude_b200.rk4_grid_tables(t.numpy(), None)
    assert np.array_equal(dt, (t[1:] - t[:-1]).numpy())
    assert list(out_step) == [-1, 0, 1, 2] and list(out_kind[1:]) == [1, 1, 1]


def test_grid_errors(built_lib):
    import ude_b200
    with pytest.raises(ude_b200.UdeError):
        ude_b200.rk4_grid_tables(np.array([0.0, 1.0, 1.0], dtype=np.float32), 0.5)     # not increasing
    with pytest.raises(ude_b200.UdeError):
        ude_b200.rk4_grid_tables(np.array([0.0, 1.0], dtype=np.float32), -1.0)         # bad step


def _desc(kind, R, L, rate, aug, B=2048):
    from ude_b200 import _lib
    d = _lib.UdeDesc()
    d.kind, d.n_regions, d.latent_dim = _lib.KIND[kind], R, L
    d.n_rate_layers = len(rate)
    for i, s in enumerate(rate):
        d.rate_sizes[i] = s
    d.n_aug_layers = len(aug)
    for i, s in enumerate(aug):
        d.aug_sizes[i] = s
    d.fa_w, d.n_groups, d.batch = 1.0, 1, B
    return d


def test_sizes_match_reference_parameter_counts(built_lib):
    """params 6 882 / 4 931 / 11 813 (R=1), 34 626 / 38 867 / 73 493 (R=49)."""
    cases = [("Fp", 1, [64, 64, 32, 2], [], 6882, 6720), ("Fa", 1, [], [64, 64, 3], 4931, 4800),
             ("FaFp", 1, [64, 64, 32, 2], [64, 64, 3], 11813, 11520),
             ("Fp", 49, [64, 64, 32, 98], [], 34626, 34368), ("FaFp", 49, [64, 64, 32, 98], [64, 64, 147], 73493, 72960)]
    for kind, R, rate, aug, n_theta, macs in cases:
        d = _desc(kind, R, 8, rate, aug)
        assert built_lib.ude_theta_size(C.byref(d)) == n_theta
        assert built_lib.ude_macs_per_eval(C.byref(d)) == macs


def test_descriptor_validation_and_plan(built_lib):
    from ude_b200 import _lib
    bad = _desc("FaFp", 1, 8, [64, 3], [64, 3])          # rate head must be 2R
    assert built_lib.ude_theta_size(C.byref(bad)) == -1
    bad = _desc("Fp", 1, 2, [8, 2], [])                  # latent_dim < 3
    assert built_lib.ude_theta_size(C.byref(bad)) == -1
    d = _desc("FaFp", 1, 8, [64, 64, 32, 2], [64, 64, 3], B=1 << 20)
    plan = _lib.UdePlan()
    assert built_lib.ude_plan(C.byref(d), 0, 8, C.byref(plan)) == 0
    assert plan.n_theta == 11813 and plan.weights_in_smem == 1
    assert plan.tape_bytes >= (1 << 20) * 8 * 4 * 3 * 4
    assert plan.smem_bytes_bwd <= 227 * 1024 and plan.smem_bytes_fwd <= 227 * 1024
    assert built_lib.ude_plan(None, 0, 8, C.byref(plan)) == -1


def test_group_plan_is_g_single_group_plans(built_lib):
    """n_groups = G (BASELINE configs[4]): the tape is G per-group blocks of the single-set plan (group-major layouts,
    include/ude_b200.h), the workspace at least that, theta size stays per group; batch must divide by G.  Small models
    run all groups in ONE launch (plan.path == 3 on the tcgen05 engine, 2 on the FFMA engine with tiles sized for the total batch)."""
    from ude_b200 import _lib
    one = _desc("Fp", 1, 8, [32, 32, 2], [], B=2048)
    p1 = _lib.UdePlan()
    assert built_lib.ude_plan(C.byref(one), _lib.METHOD["rk4"], 56, C.byref(p1)) == 0
    grp = _desc("Fp", 1, 8, [32, 32, 2], [], B=512 * 2048)
    grp.n_groups = 512
    pg = _lib.UdePlan()
    assert built_lib.ude_plan(C.byref(grp), _lib.METHOD["rk4"], 56, C.byref(pg)) == 0
    up = lambda v: (v + 255) // 256 * 256
    assert pg.workspace_bytes >= 512 * up(p1.workspace_bytes)
    assert pg.tape_bytes == 512 * up(p1.tape_bytes)
    assert pg.n_theta == p1.n_theta == built_lib.ude_theta_size(C.byref(grp))
    assert pg.path == 3 and pg.tile_traj == 128 and pg.n_ctas == 148          # default: all groups in one tcgen05 launch
    grp.flags |= 4                                                              # UDE_FLAG_FORCE_FFMA
    assert built_lib.ude_plan(C.byref(grp), _lib.METHOD["rk4"], 56, C.byref(pg)) == 0
    assert pg.path == 2 and pg.tile_traj == 64 and pg.n_ctas == 2 * 148
    grp.flags &= ~4
    grp.batch = 512 * 2048 + 1
    assert built_lib.ude_plan(C.byref(grp), _lib.METHOD["rk4"], 56, C.byref(pg)) != 0
    grp.batch = 512 * 2048
    assert built_lib.ude_plan(C.byref(grp), _lib.METHOD["dopri5"], 56, C.byref(pg)) == -2      # UDE_ERR_UNSUPPORTED


def test_plan_paths_and_modes_host_only(built_lib):
    """ude_plan is host arithmetic (148 SMs assumed without a GPU): kernel family per configuration, the Bayes and group
    modes' sizes, and the explicit rejections."""
    import ude_b200
    from ude_b200 import _lib
    L = built_lib

    def plan(mod, B, n_steps=8, kernels=None, groups=1, flags=0, method=0):
        info = ude_b200.describe_func(mod)
        info.kernels = kernels
        d = info.desc(B * groups, n_groups=groups)
        d.flags |= flags
        p = _lib.UdePlan()
        return L.ude_plan(C.byref(d), method, n_steps, C.byref(p)), p

    us = ude_b200.FaFp(1, 8, net_sizes=[64, 64, 32], aug_net_sizes=[64, 64])
    st, p = plan(us, 1)
    assert st == 0 and p.path == 1 and p.tile_traj == 128          # tcgen05 engine at every batch size
    st, p1 = plan(us, 1 << 20)
    assert st == 0 and p1.path == 1 and p1.n_ctas == 148
    assert p1.tape_bytes >= (8 * 4 + 1) * 3 * (1 << 20) * 4          # stage inputs + ONE record for the final state
    st, p = plan(us, 4096, kernels="ffma")
    assert st == 0 and p.path == 0
    st, pb = plan(us, 4096, flags=_lib.FLAG_BAYES, kernels="ffma")
    assert st == 0 and pb.path == 0 and pb.workspace_bytes > p.workspace_bytes + 2 * 32 * 11813 * 4     # weight sets + noise of 32 evaluations
    st, pt = plan(us, 4096, flags=_lib.FLAG_BAYES)
    assert st == 0 and pt.path == 1 and pt.workspace_bytes > 32 * 80_000                                # one weight image per evaluation
    assert plan(us, 4096, flags=_lib.FLAG_BAYES, method=1)[0] == -2   # Bayes funcs: rk4 only
    states = ude_b200.FaFp(49, 8, net_sizes=[64, 64, 32], aug_net_sizes=[64, 64])
    st, p = plan(states, 8192, n_steps=28)
    assert st == 0 and p.path == 0 and p.tile_traj == 32 and p.weights_in_smem == 0     # ring plan: weights stream through shared memory
    assert plan(states, 8192, kernels="tensor")[0] == -2
    assert plan(states, 2048, flags=_lib.FLAG_BAYES)[0] == -2          # too wide for the all-resident plans
    sweep = ude_b200.Fp(1, 8, net_sizes=[32, 32])
    st, p = plan(sweep, 2048, groups=64)
    assert st == 0 and p.path == 3 and p.n_ctas == 148                 # all groups in one tcgen05 launch
    st, p = plan(sweep, 2048, groups=64, kernels="ffma")
    assert st == 0 and p.path == 2


def test_front_end_argument_checks_need_no_gpu():
    """The Python front ends validate before they touch the device: wrong shapes / devices / modes raise, nothing falls back."""
    import ude_b200
    us = ude_b200.FaFp(1, 8, net_sizes=[64, 64, 32], aug_net_sizes=[64, 64])
    dec = ude_b200.Decoder(1, 8, 1)
    t = torch.arange(4.)
    with pytest.raises(RuntimeError, match="CUDA"):
        ude_b200.odeint(us, torch.zeros(4, 1, 8), t, method="rk4", options=dict(step_size=1.0))
    with pytest.raises(RuntimeError, match="CUDA"):
        ude_b200.odeint_decoded(us, dec, torch.zeros(4, 1, 8), t, options=dict(step_size=1.0))
    with pytest.raises(ValueError):
        ude_b200.odeint_decoded(us, torch.nn.Linear(4, 1), torch.zeros(4, 1, 8), t)
    with pytest.raises(ude_b200.UnsupportedFunc):
        ude_b200.odeint_decoded(us, ude_b200.Decoder(1, 8, 1, Fp=False), torch.zeros(4, 1, 8), t)
    bayes = ude_b200.Bayes_FaFp(1, 8)
    assert ude_b200.describe_func(bayes).bayes
    with pytest.raises(RuntimeError, match="CUDA"):
        ude_b200.odeint_decoded(bayes, dec, torch.zeros(4, 1, 8), t)
    with pytest.raises(ude_b200.UnsupportedFunc):
        ude_b200.describe_func(torch.nn.Linear(3, 3))
