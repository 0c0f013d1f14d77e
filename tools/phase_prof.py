"""Per-phase cycle accounting (needs a -DUDE_PROFILE build: UDE_B200_LIB=build/libude_prof.so)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ude_b200
from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
from tests.util import module_from_spec

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 64 * 4
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1
TILE = int(sys.argv[3]) if len(sys.argv) > 3 else 64
spec = UdeSpec("FaFp", R, 8, (64, 64, 32), (64, 64))
mod = module_from_spec(spec, init_theta(spec, 0), 1.0, "cuda")
y0 = synth_y0(B, R, 8, seed=1).cuda()
t = torch.arange(9.)
L = ude_b200.lib()
nw = 8
buf = (C.c_ulonglong * (nw * 32 * 2))()
NAMES = {0: "fwd L1", 1: "fwd L2", 2: "fwd L3", 3: "fwd L4(head)", 6: "fwd tail+rk", 8: "bwd dW d0 (heads)", 9: "bwd delta d0",
         10: "bwd dW d1", 11: "bwd delta d1", 12: "bwd dW d2", 13: "bwd delta d2", 14: "bwd dW d3", 15: "bwd delta d3",
         20: "bwd tail vjp", 21: "bwd combine"}


def dump(title, n_evals):
    L.ude_debug_read_prof(buf)
    a = np.array(list(buf), dtype=np.float64).reshape(nw, 32, 2)
    print(f"== {title}: cycles per tile-eval (CTA 0; {n_evals} tile-evals)")
    tot = 0
    for pid, name in NAMES.items():
        work, wait = a[:, pid, 0] / n_evals, a[:, pid, 1] / n_evals
        if work.max() == 0 and wait.max() == 0:
            continue
        span = (work + wait).mean()
        tot += span
        print(f"  {name:20s} span {span:8.0f}  work/warp min {work.min():7.0f} mean {work.mean():7.0f} max {work.max():7.0f}")
    print(f"  sum of phase spans {tot:8.0f}")


tiles_cta0 = (B // TILE + 147) // 148
y0g = y0.clone().requires_grad_(True)
y = ude_b200.odeint(mod, y0g, t, method="rk4", options=dict(step_size=1.0, kernels="ffma"))
torch.cuda.synchronize()
dump("forward kernel", tiles_cta0 * 32)
(y[..., :3] ** 2).sum().backward()
torch.cuda.synchronize()
dump("backward kernel", tiles_cta0 * 32)
