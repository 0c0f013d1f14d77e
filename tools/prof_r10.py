"""hhs regional model (R = 10, run_ode.py:49-57) timing: FaFp(10, 8, [64,64,32], [64,64]), fwd + adjoint per iteration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ude_b200
from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
from tests.util import module_from_spec
spec = UdeSpec("FaFp", 10, 8, (64, 64, 32), (64, 64))
mod = module_from_spec(spec, init_theta(spec, 0), 1.0, "cuda")
import ctypes as C
from ude_b200 import _lib
for B, steps in ((2048, 7), (2048, 28), (65536, 8), (262144, 8)):
    t = torch.arange(steps + 1.) if steps < 20 else torch.arange(steps + 1.) / 7
    y0 = synth_y0(B, 10, 8, seed=1).cuda()
    def step():
        y0g = y0.clone().requires_grad_(True); mod.zero_grad(set_to_none=True); mod.clear_tracking()
        y = ude_b200.odeint(mod, y0g, t, method="rk4", options=dict(step_size=t[1] - t[0])); (y[..., 1] ** 2).mean().backward()
    for _ in range(2): step()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    p = _lib.UdePlan(); d = ude_b200.describe_func(mod).desc(B)
    ude_b200.lib().ude_plan(C.byref(d), 0, steps, C.byref(p))
    ts = B * steps / (ms * 1e-3)
    print(f"R=10 B={B} steps={steps}: {ms:.2f} ms, {ts:.3e} traj-steps/s, frac_ffma {ts * 24 * spec.macs_per_eval() / 72.4e12:.3f}; plan tile {p.tile_traj} ctas {p.n_ctas} w_smem {p.weights_in_smem} smem {p.smem_bytes_fwd}/{p.smem_bytes_bwd}")
