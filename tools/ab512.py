"""Timing probes of the R = 49 shapes (device time per fwd + adjoint iteration)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ude_b200
from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
from tests.util import module_from_spec
spec = UdeSpec("FaFp", 49, 8, (64, 64, 32), (64, 64))
mod = module_from_spec(spec, init_theta(spec, 0), 1.0, "cuda")
def run(method, B, t, seed=1, clear=True, reps=3, **kw):
    y0 = synth_y0(B, 49, 8, seed=seed).cuda()
    def step():
        y0g = y0.clone().requires_grad_(True); mod.zero_grad(set_to_none=True)
        if clear: mod.clear_tracking()
        y = ude_b200.odeint(mod, y0g, t, method=method, **kw); (y[..., 1] ** 2).mean().backward()
    for _ in range(2): step()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): step()
    e1.record(); torch.cuda.synchronize()
    print(f"{method} B={B} seed={seed} clear={clear}: {e0.elapsed_time(e1)/reps:.2f} ms", kw.get("options", {}).get("per_trajectory"), flush=True)
t = torch.arange(29.) / 7
for which in sys.argv[1:] or ["d1", "d2", "d2c", "r256", "r256b", "r1024"]:
    if which == "d1": run("dopri5", 2048, torch.arange(9.), seed=1, rtol=1e-5, atol=1e-9, options=dict(per_trajectory=True))
    if which == "d2": run("dopri5", 2048, torch.arange(9.), seed=2, rtol=1e-5, atol=1e-9, options=dict(per_trajectory=True))
    if which == "d2c": run("dopri5", 2048, torch.arange(9.), seed=2, clear=False, rtol=1e-5, atol=1e-9, options=dict(per_trajectory=True))
    if which == "r256": run("rk4", 256, t, options=dict(step_size=t[1] - t[0]))
    if which == "r256b": run("rk4", 256, t, reps=10, options=dict(step_size=t[1] - t[0]))
    if which == "r1024": run("rk4", 1024, t, options=dict(step_size=t[1] - t[0]))
