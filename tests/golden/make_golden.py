"""Generate the committed golden fixtures (run in the BUILD container only).

Imports the REAL reference ODE funcs from /root/reference/lib/models.py
(Fp :109-156, Fa :158-197, FaFp :199-265), drives them with the oracle's
restated torchdiffeq solver (oracle/odeint_ref.py; torchdiffeq itself is absent
offline) and stores inputs + outputs + gradients as compressed .npz.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

The loss used for the gradients is linear in the outputs and in the five
side-channel sums so that the cotangents handed to the fused backward are
exactly ``G`` and ``c_side``:

    L = sum(y * G) + c_side . [sum beta, sum gamma, sum beta^2, sum gamma^2, sum Fa^2]
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import lib.models as ref_models  # noqa: E402  (the real reference classes)

from oracle import odeint_ref  # noqa: E402
from oracle.ude_rhs import AUG_HEAD_SCALE, synth_y0, theta_from_module  # noqa: E402

C_SIDE = torch.tensor([0.3, -0.2, 0.05, 0.07, 0.11], dtype=torch.float64)


def build(kind, R, L, seed, fa_w=1.0, **kw):
    torch.manual_seed(seed)
    mod = getattr(ref_models, kind)(R, L, **kw)
    with torch.no_grad():
        if kind in ("Fa", "FaFp"):       # keep trajectories inside the clamp range
            mod.aug_net[-1].weight.mul_(AUG_HEAD_SCALE)
            mod.aug_net[-1].bias.mul_(AUG_HEAD_SCALE)
    if kind == "FaFp":
        mod.Fa_w = fa_w
    return mod


def module_side_sums(mod):
    out = torch.zeros(6, dtype=torch.float64)
    if getattr(mod, "params", None):
        p = torch.cat([q.reshape(-1, 2) for q in mod.params]).double()
        out[0:2] = p.sum(0)
        out[2:4] = (p * p).sum(0)
        out[5] = p.shape[0]
    if getattr(mod, "tracker", None):
        fa = torch.cat([q.reshape(-1, 3) for q in mod.tracker]).double()
        out[4] = (fa * fa).sum()
        if out[5] == 0:
            out[5] = fa.shape[0]
    return out


def run_case(name, mod, y0, t, method, options, rtol=1e-7, atol=1e-9, seed=0):
    spec, theta, fa_w = theta_from_module(mod)
    mod.clear_tracking()
    y0 = y0.clone().requires_grad_(True)
    stats = None
    opts = dict(options)
    h = opts.get("step_size", None)
    if method == "dopri5":
        stats = odeint_ref.Dopri5Stats()
        opts["stats"] = stats
    y = odeint_ref.odeint(mod, y0, t, rtol=rtol, atol=atol, method=method, options=opts)
    side = module_side_sums(mod)
    g = torch.Generator().manual_seed(1000 + seed)
    G = torch.randn(y.shape, generator=g, dtype=torch.float64).to(y.dtype)
    loss = (y.double() * G.double()).sum() + (side[:5] * C_SIDE).sum()
    params = [p for p in mod.parameters()]
    grads = torch.autograd.grad(loss, [y0, *params])
    # flatten parameter grads in theta order (ModuleList order == parameters() order)
    dtheta = torch.cat([gr.reshape(-1) for gr in grads[1:]])
    assert dtheta.numel() == theta.numel()
    out = dict(
        kind=spec.kind, n_regions=spec.n_regions, latent_dim=spec.latent_dim,
        net_sizes=np.array(spec.net_sizes, dtype=np.int64),
        aug_sizes=np.array(spec.aug_sizes, dtype=np.int64),
        fa_w=np.float64(fa_w), method=method,
        step_size=np.float32(h if h is not None else np.nan),
        rtol=np.float64(rtol), atol=np.float64(atol),
        theta=theta.numpy(), y0=y0.detach().numpy(), t=t.numpy(),
        y=y.detach().numpy(), G=G.numpy(), c_side=C_SIDE.numpy(),
        side=side.detach().numpy(), dy0=grads[0].numpy(), dtheta=dtheta.numpy(),
    )
    out["first_step"] = np.float64(opts.get("first_step", -1.0) or -1.0)
    if stats is not None:
        out["n_accept"] = stats.n_accept.numpy()
        out["n_reject"] = stats.n_reject.numpy()
        out["dt0"] = stats.dt0.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: y{tuple(y.shape)} theta {theta.numel()} side {side.tolist()}")


def main():
    big = dict(net_sizes=[64, 64, 32], aug_net_sizes=[64, 64])
    # 1. headline shape, on-grid weekly steps (run_ode.py:93,147; lib/VAE.py:126)
    m = build("FaFp", 1, 8, 1, **big)
    run_case("rk4_fafp_r1_weekly", m, synth_y0(7, 1, 8, seed=11), torch.arange(9.), "rk4",
             dict(step_size=1.0), seed=1)
    # 2. daily grid, h = fl32(1/7): ulp-mismatch -> lerp branch (run_ode.py:130)
    m = build("Fp", 1, 8, 2, net_sizes=[64, 64, 32])
    t = torch.arange(57.) / 7
    run_case("rk4_fp_r1_daily", m, synth_y0(5, 1, 8, seed=12), t, "rk4",
             dict(step_size=(t[1] - t[0])), seed=2)
    # 3. off-grid outputs, short last step (tuning/tune_node.py:192,204)
    m = build("Fa", 1, 8, 3, aug_net_sizes=[64, 64])
    t = torch.linspace(1, 56, 56) / 7
    run_case("rk4_fa_r1_interp", m, synth_y0(4, 1, 8, seed=13), t, "rk4", dict(step_size=1.0), seed=3)
    # 4. coupled regions, Fa_w != 1
    m = build("FaFp", 10, 8, 4, fa_w=0.3, **big)
    run_case("rk4_fafp_r10", m, synth_y0(3, 10, 8, seed=14), torch.arange(5.), "rk4",
             dict(step_size=1.0), seed=4)
    # 5. class-default nets, L=4, some states outside [-1, 2] (clamp)
    m = build("Fp", 1, 4, 5)
    y0 = synth_y0(6, 1, 4, seed=15)
    y0[0, 0, 0] = 2.5
    y0[1, 0, 1] = -1.25
    y0[2, 0, 2] = 2.75
    run_case("rk4_fp_default_clamp", m, y0, torch.arange(4.), "rk4", dict(step_size=0.5), seed=5)
    # 6. 49 regions (run_ode.py:60)
    m = build("FaFp", 49, 8, 6, **big)
    run_case("rk4_fafp_r49", m, synth_y0(2, 49, 8, seed=16), torch.arange(3.), "rk4",
             dict(step_size=1.0), seed=6)
    # 7. L=16, Fa default nets
    m = build("Fa", 2, 16, 7)
    run_case("rk4_fa_r2_l16", m, synth_y0(5, 2, 16, seed=17), torch.arange(4.), "rk4",
             dict(step_size=1.0), seed=7)
    # 8/9. Dopri5, per-trajectory control (BASELINE config 2 tolerances)
    m = build("FaFp", 1, 8, 8, **big)
    run_case("dopri5_fafp_r1", m, synth_y0(6, 1, 8, seed=18), torch.arange(9.), "dopri5",
             dict(per_trajectory=True), rtol=1e-5, atol=1e-9, seed=8)
    m = build("Fp", 3, 8, 9, net_sizes=[32, 32])
    run_case("dopri5_fp_r3", m, synth_y0(4, 3, 8, seed=19), torch.arange(0., 4.5, 0.5), "dopri5",
             dict(per_trajectory=True), rtol=1e-6, atol=1e-8, seed=9)
    # 10/11. the same machinery with a DETERMINISTIC controller: tolerance so loose that every step
    # is accepted and grows by ifactor exactly, so fp32 round-off in the error norm cannot change the
    # step sequence (at tight tolerances it does, in any two fp32 implementations)
    m = build("FaFp", 1, 8, 10, **big)
    run_case("dopri5det_fafp_r1", m, synth_y0(70, 1, 8, seed=20), torch.arange(9.), "dopri5",
             dict(per_trajectory=True, first_step=0.01), rtol=1e-1, atol=1e-1, seed=10)
    m = build("FaFp", 3, 8, 11, fa_w=0.5, net_sizes=[32, 32], aug_net_sizes=[20])
    run_case("dopri5det_fafp_r3", m, synth_y0(5, 3, 8, seed=21), torch.arange(0., 4.5, 0.5), "dopri5",
             dict(per_trajectory=True, first_step=0.05), rtol=1e-1, atol=1e-1, seed=11)


if __name__ == "__main__":
    main()
