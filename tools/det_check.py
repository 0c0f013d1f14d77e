"""Bit-reproducibility check of the tensor kernel family: 30 reruns of one fused forward + adjoint (B = 40 000) must
equal the first run bit for bit.  python tools/det_check.py  (GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ude_b200
from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
from tests.util import module_from_spec, flat_grad
spec = UdeSpec("FaFp", 1, 8, (64, 64, 32), (64, 64))
theta = init_theta(spec, seed=7)
B = 40000
y0 = synth_y0(B, 1, 8, seed=3)
t = torch.arange(5.)
G = torch.randn(5, B, 1, 8, generator=torch.Generator().manual_seed(6))
def run():
    mod = module_from_spec(spec, theta, 1.0, "cuda")
    y0g = y0.cuda().requires_grad_(True)
    y = ude_b200.odeint(mod, y0g, t, method="rk4", options=dict(step_size=1.0, kernels="tensor"))
    (y * G.cuda()).sum().backward()
    return y.detach(), y0g.grad, flat_grad(mod)
ref = run()
bad = [0, 0, 0]
for i in range(30):
    r = run()
    for k in range(3):
        if not torch.equal(r[k], ref[k]):
            bad[k] += 1
            if k == 2 and bad[k] <= 3:
                d = (r[k] - ref[k]).abs()
                print("dtheta differs: n=", int((d > 0).sum()), "max", float(d.max()), "rel", float(d.max() / ref[k].abs().max()), "idx", d.argmax().item())
print("mismatches (y, dy0, dtheta) in 30 runs:", bad)
