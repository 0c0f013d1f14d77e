"""Bayes_FaFp (R = 1, the reference's newest runs: run_ode.ipynb cell 3) fwd + adjoint timing on both engines."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ude_b200
from oracle.ude_rhs import synth_y0
torch.manual_seed(0)
mod = ude_b200.Bayes_FaFp(1, 8, net_sizes=[64, 64, 32], aug_net_sizes=[64, 64], prior_std=0.05).cuda()
with torch.no_grad():
    for p in (mod.aug_net[-1].w_mean, mod.aug_net[-1].b_mean):
        p.mul_(0.05)
for B, steps in ((2048, 7), (1 << 17, 8), (1 << 20, 8)):
    t = torch.arange(steps + 1.)
    y0 = synth_y0(B, 1, 8, seed=1).cuda()
    for kernels in ("ffma", "tensor"):
        if kernels == "ffma" and B > (1 << 17):
            continue
        def step():
            y0g = y0.clone().requires_grad_(True); mod.zero_grad(set_to_none=True); mod.clear_tracking()
            y = ude_b200.odeint(mod, y0g, t, method="rk4", options=dict(step_size=1.0, kernels=kernels)); (y[..., 1] ** 2).mean().backward()
        for _ in range(2): step()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"Bayes_FaFp R=1 B={B} steps={steps} {kernels}: {ms:.2f} ms, {B * steps / (ms * 1e-3):.3e} traj-steps/s", flush=True)
