"""FFMA engine vs tcgen05 engine on the R = 1 headline model at small batches: device time per fwd + adjoint iteration,
eager (host-bound below ~1 ms) and replayed from a CUDA graph (ude_b200.GraphedIteration, no optimizer).
usage (GPU box): python tools/engine_crossover.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ude_b200
from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
from tests.util import module_from_spec

spec = UdeSpec("FaFp", 1, 8, (64, 64, 32), (64, 64))
mod = module_from_spec(spec, init_theta(spec, 0), 1.0, "cuda")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for steps in (7, 85):
    t = torch.arange(steps + 1.) if steps == 7 else torch.arange(steps + 1.) / 7
    for B in (1, 128, 2048, 4096, 8192, 18944):
        y0 = synth_y0(B, 1, 8, seed=1).cuda().requires_grad_(True)
        row = []
        for kernels in ("ffma", "tensor"):
            mod_g = module_from_spec(spec, init_theta(spec, 0), 1.0, "cuda")       # own module: its AccumulateGrad nodes live on the capture stream

            def step_fn(y, m=mod):
                out = ude_b200.odeint(m, y, t, method="rk4", options=dict(step_size=t[1] - t[0], kernels=kernels))
                return (out[..., 1] ** 2).mean()

            def eager():
                mod.zero_grad(set_to_none=True)
                step_fn(y0).backward()
            it = ude_b200.GraphedIteration(lambda y: step_fn(y, mod_g), list(mod_g.parameters()))
            y0_g = y0.detach().clone().requires_grad_(True)
            row.append((timed(eager, 10), timed(lambda: it(y0_g))))
        print(f"steps={steps} B={B}: ffma eager {row[0][0]:.3f} graph {row[0][1]:.3f} ms | tensor eager {row[1][0]:.3f} graph {row[1][1]:.3f} ms", flush=True)
