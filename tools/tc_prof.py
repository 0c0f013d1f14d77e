"""Per-category cycle breakdown of the tensor-core kernels (CTA 0), from a -DUDE_TC_PROFILE build.
  build (here):  python tools/tc_prof.py build
  run (GPU box): UDE_B200_LIB=$PWD/build/libude_tcprof.so python tools/tc_prof.py [B]"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "forecasting-influenza-using-universal-differential-equations_b200", "csrc")
LIB = os.path.join(ROOT, "build", "libude_tcprof.so")

if len(sys.argv) > 1 and sys.argv[1] == "build":
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off"]
    obj = os.path.join(ROOT, "build", "ude_tc_host_prof.o")
    subprocess.check_call(["nvcc", *flags, "-DUDE_TC_PROFILE", "-c", "-o", obj, os.path.join(CSRC, "ude_tc_host.cu")])
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB,
                           os.path.join(ROOT, "build", "ude_host.o"), obj])
    print("built", LIB)
    sys.exit(0)

import torch

import ude_b200
from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
from tests.util import module_from_spec

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 2
spec = UdeSpec("FaFp", 1, 8, (64, 64, 32), (64, 64))
mod = module_from_spec(spec, init_theta(spec, 0), 1.0, "cuda")
y0 = synth_y0(B, 1, 8, seed=1).cuda()
t = torch.arange(9.)
L = ude_b200.lib()
names = ["sync", "image+elu'+stepH", "wait", "epi_fwd", "delta_epi", "flush+dx", "total", "-"]


def dump(tag):
    buf = (C.c_ulonglong * 16)()
    L.ude_debug_tc_prof(buf)
    for th, off in (("thread 0  ", 0), ("thread 128", 8)):
        tot = buf[off + 6]
        parts = "  ".join(f"{names[i]} {100.0 * buf[off + i] / tot:5.1f}%" for i in range(6))
        print(f"{tag} {th}: total {tot / 1e6:8.2f} Mcycles | {parts}")
    evals = 2 * 8 * 4       # tiles of CTA 0 x steps x stages
    print(f"{tag} cycles per tile-eval: {buf[6] / evals:9.0f}")


for it in range(2):
    y0g = y0.clone().requires_grad_(True)
    y = ude_b200.odeint(mod, y0g, t, method="rk4", options=dict(step_size=1.0, kernels="tensor"))
    torch.cuda.synchronize()
    if it == 1:
        dump("fwd")
    (y[..., :3] ** 2).sum().backward()
    torch.cuda.synchronize()
    if it == 1:
        dump("bwd")
