"""Secondary timings of the other BASELINE.json configs (NOT the bench line: bench.py measures configs[3]).
  python tools/bench_configs.py            -> one JSON line per config
configs[1] 49-state Dopri5 epoch, configs[2] 49 states x 256 samples RK4 28 steps, configs[4] 64 weight-set groups per GPU."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ude_b200
from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
from tests.util import module_from_spec


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def macs(spec):
    D = spec.n_regions * spec.latent_dim
    m = 0
    for sizes, head in ((spec.net_sizes, 2 * spec.n_regions), (spec.aug_sizes, 3 * spec.n_regions)):
        if len(sizes):
            dims = [D, *sizes, head]
            m += sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    return m


peak = 72.4e12      # measured FFMA peak (bench.py: ude_measure_ffma_tflops)
out = []

# ---- configs[2]: 49 states x 256 latent samples x b = 32, RK4, 28 daily steps, fwd + backprop
spec = UdeSpec("FaFp", 49, 8, (64, 64, 32), (64, 64))
mod = module_from_spec(spec, init_theta(spec, 0), 1.0, "cuda")
for b in (1, 32):
    B = 256 * b
    y0 = synth_y0(B, 49, 8, seed=1).cuda()
    t = torch.arange(29.) / 7

    def step():
        y0g = y0.clone().requires_grad_(True)
        mod.zero_grad(set_to_none=True)
        y = ude_b200.odeint(mod, y0g, t, method="rk4", options=dict(step_size=t[1] - t[0]))
        (y[..., 1] ** 2).mean().backward()

    ms = timed(step)
    ts = B * 28 / (ms * 1e-3)
    out.append({"config": f"configs[2] FaFp R=49, y0 [{B},49,8], RK4 28 daily steps, fwd+adjoint (FFMA engine)", "ms_per_iter": ms,
                "trajectory_steps_per_s": ts, "frac_of_ffma_peak": ts * 24 * macs(spec) / peak})

# ---- configs[1]: 49-state regional batch, Dopri5 rtol 1e-5, one epoch = 153 iterations of [2048,49,8]
y0 = synth_y0(2048, 49, 8, seed=2).cuda()
t = torch.arange(9.)


def step_d5():
    y0g = y0.clone().requires_grad_(True)
    mod.zero_grad(set_to_none=True)
    y = ude_b200.odeint(mod, y0g, t, rtol=1e-5, atol=1e-9, method="dopri5")
    (y[..., 1] ** 2).mean().backward()


ms = timed(step_d5)
out.append({"config": "configs[1] FaFp R=49, y0 [2048,49,8], Dopri5 rtol 1e-5 atol 1e-9, t=[0..8], fwd+adjoint (FFMA engine)",
            "ms_per_iter": ms, "epoch_153_iters_s": ms * 153e-3})

# ---- configs[4]: 64 groups per GPU (512 over 8), Fp hidden 32, 2048 trajectories each, 56 outputs on a weekly grid
spec = UdeSpec("Fp", 1, 8, (32, 32), ())
G, Bg = 64, 2048
mods = [module_from_spec(spec, init_theta(spec, g), 1.0, "cuda") for g in range(G)]
y0 = synth_y0(G * Bg, 1, 8, seed=3).reshape(G, Bg, 1, 8).cuda()
t = torch.linspace(1, 56, 56) / 7


def step_g():
    y0g = y0.clone().requires_grad_(True)
    for m in mods:
        m.zero_grad(set_to_none=True)
    y = ude_b200.odeint_groups(mods, y0g, t, method="rk4", options=dict(step_size=1.0))
    (y[..., 1] ** 2).mean().backward()


ms = timed(step_g)
n_steps = 8
ts = G * Bg * n_steps / (ms * 1e-3)
out.append({"config": f"configs[4] {G} groups x Fp(R=1,L=8,[32,32]) x 2048 trajectories, RK4 h=1 over t=[1/7..8] (8 grid steps, 56 outputs), "
                      "fwd+adjoint, stream fan-out (FFMA engine)", "ms_per_iter": ms, "trajectory_steps_per_s": ts,
            "frac_of_ffma_peak": ts * 24 * macs(spec) / peak})


def seq():
    for g in range(G):
        y0g = y0[g].clone().requires_grad_(True)
        y = ude_b200.odeint(mods[g], y0g, t, method="rk4", options=dict(step_size=1.0))
        (y[..., 1] ** 2).mean().backward()


ms_seq = timed(seq, reps=2, warm=1)
out[-1]["ms_per_iter_as_64_separate_odeint_calls"] = ms_seq
for o in out:
    print(json.dumps(o))
