"""A/B of the ring plan (UDE_B200_NO_RING=1 switches it off): R = 49 RK4 and Dopri5, forward and backward timed apart."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ude_b200
from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
from tests.util import module_from_spec, flat_grad

spec = UdeSpec("FaFp", 49, 8, (64, 64, 32), (64, 64))
mod = module_from_spec(spec, init_theta(spec, 0), 1.0, "cuda")

def run(method, B, t, **kw):
    y0 = synth_y0(B, 49, 8, seed=1).cuda()
    res = []
    for it in range(4):
        y0g = y0.clone().requires_grad_(True)
        mod.zero_grad(set_to_none=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        y = ude_b200.odeint(mod, y0g, t, method=method, **kw)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        (y[..., 1] ** 2).mean().backward()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        res.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    g = flat_grad(mod)
    print(f"{method} B={B} ring={'off' if os.environ.get('UDE_B200_NO_RING') else 'on'}: fwd {res[-1][0]:.2f} ms bwd {res[-1][1]:.2f} ms; "
          f"finite y {bool(torch.isfinite(y).all())} dy0 {bool(torch.isfinite(y0g.grad).all())} dtheta {bool(torch.isfinite(g).all())} "
          f"|dtheta| {float(g.abs().max()):.4e} sum y {float(y.double().sum()):.6f}")

t = torch.arange(29.) / 7
run("rk4", 8192, t, options=dict(step_size=t[1] - t[0]))
run("dopri5", 2048, torch.arange(9.), rtol=1e-5, atol=1e-9, options=dict(per_trajectory=True))
run("dopri5", 12, torch.arange(9.), rtol=1e-5, atol=1e-9, options=dict(per_trajectory=True))
