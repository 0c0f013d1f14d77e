"""Quick parity check of the tcgen05 path against the oracle and the FFMA engine:
python tools/tc_check.py [B] [n_t] [kind] [fwd|all]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ude_b200
from oracle import odeint_ref
from oracle.ude_rhs import OracleFunc, UdeSpec, init_theta, synth_y0
from tests.util import flat_grad, module_from_spec, rel_err

B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
n_t = int(sys.argv[2]) if len(sys.argv) > 2 else 4
kind = sys.argv[3] if len(sys.argv) > 3 else "FaFp"
what = sys.argv[4] if len(sys.argv) > 4 else "all"
spec = UdeSpec(kind, 1, 8, (64, 64, 32) if kind != "Fa" else (), (64, 64) if kind != "Fp" else ())
theta = init_theta(spec, seed=5)
y0 = synth_y0(B, 1, 8, seed=B)
t = torch.arange(float(n_t))
G = torch.randn(n_t, B, 1, 8, generator=torch.Generator().manual_seed(2))

ref_f = OracleFunc(spec, theta)
y0r = y0.clone().requires_grad_(True)
yr = odeint_ref.odeint(ref_f, y0r, t, method="rk4", options=dict(step_size=1.0))
(yr * G).sum().backward()

res = {}
for kern in ("ffma", "tensor"):
    mod = module_from_spec(spec, theta, 1.0, "cuda")
    y0g = y0.cuda().requires_grad_(True)
    mod.clear_tracking()
    y = ude_b200.odeint(mod, y0g, t, method="rk4", options=dict(step_size=1.0, kernels=kern))
    torch.cuda.synchronize()
    print(f"[{kern}] y   rel err vs oracle {rel_err(y, yr):.3e}", flush=True)
    if what == "all":
        (y * G.cuda()).sum().backward()
        torch.cuda.synchronize()
        dth = flat_grad(mod)
        print(f"[{kern}] dy0 rel err vs oracle {rel_err(y0g.grad, y0r.grad):.3e}")
        print(f"[{kern}] dth rel err vs oracle {rel_err(dth, ref_f.theta.grad):.3e}", flush=True)
        # per layer
        off = 0
        for net, n_in, n_out in spec.layer_dims():
            for nm, cnt in (("W", n_in * n_out), ("b", n_out)):
                a = dth[off:off + cnt].cpu().double()
                b = ref_f.theta.grad[off:off + cnt].double()
                print(f"    {net} {n_in}x{n_out} {nm}: max|ref| {float(b.abs().max()):.3e} max err {float((a - b).abs().max()):.3e}")
                off += cnt
        res[kern] = (y.detach(), y0g.grad.detach(), dth.detach())
    else:
        res[kern] = (y.detach(),)
print("tensor vs ffma:", [f"{rel_err(a, b):.3e}" for a, b in zip(res["tensor"], res["ffma"])])
