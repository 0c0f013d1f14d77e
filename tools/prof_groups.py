"""configs[4] iteration (64 weight-set groups x Fp(R=1,L=8,[32,32]) x 2048 trajectories) for launch lists / timing.
usage: python tools/prof_groups.py [G] [flat]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ude_b200
from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
from tests.util import module_from_spec

G = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Bg = 2048
spec = UdeSpec("Fp", 1, 8, (32, 32), ())
mods = [module_from_spec(spec, init_theta(spec, g), 1.0, "cuda") for g in range(G)]
y0 = synth_y0(G * Bg, 1, 8, seed=3).reshape(G, Bg, 1, 8).cuda()
t = torch.linspace(1, 56, 56) / 7


def step():
    y0g = y0.clone().requires_grad_(True)
    for m in mods:
        m.zero_grad(set_to_none=True)
    y = ude_b200.odeint_groups(mods, y0g, t, method="rk4", options=dict(step_size=1.0))
    (y[..., 1] ** 2).mean().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
print(f"G={G}: {e0.elapsed_time(e1) / 5:.3f} ms per iteration")

# ---- the same sweep with the weights of all G models in ONE [G, n_theta] leaf (flatten_group_parameters) and ONE Adam
flat = ude_b200.flatten_group_parameters(mods)
opt = torch.optim.Adam([flat], lr=1e-3, capturable=True)


def step_flat():
    flat.grad = None
    y = ude_b200.odeint_groups(mods, y0, t, method="rk4", options=dict(step_size=1.0), theta=flat)
    (y[..., 1] ** 2).mean().backward()


for _ in range(3):
    step_flat()
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    step_flat()
e1.record()
torch.cuda.synchronize()
print(f"G={G}, flat theta: {e0.elapsed_time(e1) / 5:.3f} ms per iteration (eager)")


def step_fn(y):
    yy = ude_b200.odeint_groups(mods, y, t, method="rk4", options=dict(step_size=1.0), theta=flat)
    return (yy[..., 1] ** 2).mean()


# fresh objects for the captured variant (parameters that were trained eagerly keep their gradient-accumulator nodes on the default stream)
mods = [module_from_spec(spec, init_theta(spec, g), 1.0, "cuda") for g in range(G)]
flat = ude_b200.flatten_group_parameters(mods)
opt = torch.optim.Adam([flat], lr=1e-3, capturable=True)
it = ude_b200.GraphedIteration(step_fn, [flat], opt, grad_lim=None)
for _ in range(3):
    it(y0)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    it(y0)
e1.record()
torch.cuda.synchronize()
print(f"G={G}, flat theta, CUDA graph incl. Adam: {e0.elapsed_time(e1) / 10:.3f} ms per iteration")
