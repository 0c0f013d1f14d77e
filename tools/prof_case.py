"""Small fused fwd+bwd run for ncu captures: python tools/prof_case.py [B] [kind] [R]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ude_b200
from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
from tests.util import module_from_spec

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 64 * 2
kind = sys.argv[2] if len(sys.argv) > 2 else "FaFp"
R = int(sys.argv[3]) if len(sys.argv) > 3 else 1
method = sys.argv[4] if len(sys.argv) > 4 else "rk4"
spec = UdeSpec(kind, R, 8, (64, 64, 32) if kind != "Fa" else (), (64, 64) if kind != "Fp" else ())
mod = module_from_spec(spec, init_theta(spec, 0), 1.0, "cuda")
y0 = synth_y0(B, R, 8, seed=1).cuda()
t = torch.arange(9.)
for it in range(3):
    y0g = y0.clone().requires_grad_(True)
    mod.clear_tracking()
    if method == "rk4":
        y = ude_b200.odeint(mod, y0g, t, method="rk4", options=dict(step_size=1.0, kernels=os.environ.get("UDE_KERNELS") or None))
    else:
        y = ude_b200.odeint(mod, y0g, t, rtol=1e-5, atol=1e-9, method="dopri5")
    (y[..., :3] ** 2).sum().backward()
torch.cuda.synchronize()
print("ok", float(y.sum()))
