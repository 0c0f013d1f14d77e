"""One fused-consumer iteration at the bench size for ncu captures: ude_b200.odeint_decoded (f2) forward + adjoint on the
tcgen05 engine, B = 2**20, 8 weekly steps.  usage: python tools/prof_decoded.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ude_b200
from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
from tests.util import module_from_spec

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
spec = UdeSpec("FaFp", 1, 8, (64, 64, 32), (64, 64))
mod = module_from_spec(spec, init_theta(spec, 0), 1.0, "cuda")
dec = ude_b200.Decoder(1, 8, 1).cuda()
y0 = synth_y0(B, 1, 8, seed=1).cuda()
t = torch.arange(9.)
for it in range(3):
    y0g = y0.clone().requires_grad_(True)
    mod.clear_tracking()
    yp, reg = ude_b200.odeint_decoded(mod, dec, y0g, t, options=dict(step_size=1.0))
    ((yp - 0.1) ** 2).mean().add(0.1 * reg / B).backward()
torch.cuda.synchronize()
print("ok", float(yp.sum()), float(reg))
