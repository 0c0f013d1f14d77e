"""Where the end-to-end step of bench.py spends its time (1 GPU, B = 2**20): CUDA events around H2D / solve / loss / backward / D2H."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import ude_b200
from oracle.ude_rhs import init_theta, synth_y0

dev = torch.device("cuda", 0)
spec = bench._spec()
mod = bench._module(spec, init_theta(spec, 0), dev)
dec = bench._decoder(dev)
t = torch.arange(9.)
y_ili = (0.02 + 0.5 * torch.exp(-((t - 5.0) / 4.0) ** 2)).reshape(-1, 1, 1).to(dev)
B = 1 << 20
y0_host = synth_y0(B, 1, 8, seed=100).pin_memory()
names = ["h2d", "solve_fwd", "loss", "backward", "d2h"]


def step(ev):
    ev[0].record()
    y0d = y0_host.to(dev, non_blocking=True).requires_grad_(True)
    mod.clear_tracking(); mod.zero_grad(set_to_none=True); dec.zero_grad(set_to_none=True)
    ev[1].record()
    yp, reg = ude_b200.odeint_decoded(mod, dec, y0d, t, options=dict(step_size=t[1] - t[0]))
    ev[2].record()
    ls = bench._loss_tail(yp, y_ili, reg, yp.shape[1], mod)
    ev[3].record()
    ls.backward()
    ev[4].record()
    flat = torch.cat([p.grad.reshape(-1) for p in (*mod.parameters(), *dec.parameters())])
    out = torch.cat([ls.detach().reshape(1), flat]).to("cpu")
    ev[5].record()
    return out


for _ in range(2):
    step([torch.cuda.Event(enable_timing=True) for _ in range(6)])
torch.cuda.synchronize()
acc = [0.0] * 5
wall = 0.0
for _ in range(5):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    t0 = time.perf_counter()
    step(ev)
    torch.cuda.synchronize()
    wall += time.perf_counter() - t0
    for i in range(5):
        acc[i] += ev[i].elapsed_time(ev[i + 1])
print("per step (ms):", {n: round(a / 5, 3) for n, a in zip(names, acc)}, "sum", round(sum(acc) / 5, 3), "wall", round(wall / 5 * 1e3, 3))
