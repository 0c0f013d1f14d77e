#!/usr/bin/env python
"""bench.py — UDE trajectory-steps/s (forward + discrete adjoint) on N B200s.

Contract (driver): ``python bench.py --gpus N --steps K --warmup W`` (torchrun for N>1)
prints ONE JSON line on rank 0.  A "step" is one pass of the hot path over one batch of
synthetic ILI-shaped input: the fused RK4 forward over the whole time grid followed by the
fused discrete-adjoint backward (+ the NCCL all-reduce of dL/dtheta when N>1).

Workload = BASELINE.json configs[3] (the configuration the metric is quoted on):
  FaFp(R=1, L=8, net [64,64,32], aug [64,64]), B = 2**20 trajectories PER GPU (weak scaling:
  trajectories are independent IVPs, sharded across ranks), t = [0..8] weeks, h = 1 (8 RK4 steps).

``--impl reference`` times the reference's CPU path for the same metric/config on the host
cores: the oracle port (oracle/odeint_ref.py driving the restated lib/models.py RHS with
torch autograd — torchdiffeq itself is not installable offline), on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_STEPS_GRID = 8            # weekly RK4 steps, t = arange(9.)  (run_ode.py:93,147)
B_PER_GPU = 1 << 20
CPU_SAMPLE_B = 65536
METRIC = "ude_trajectory_steps_per_sec_fwd_plus_adjoint"
UNIT = "trajectory-steps/s"


def _spec():
    from oracle.ude_rhs import UdeSpec
    return UdeSpec("FaFp", 1, 8, (64, 64, 32), (64, 64))


def tensor_macs_per_tile_eval(spec) -> dict:
    """bf16 MACs the tcgen05 kernels ISSUE per 128-trajectory tile and RHS evaluation (3-way bf16 split: six
    products per fp32 product in the layer chain and the delta chain, three first-order products in the
    weight-gradient contraction; dims padded to 16) -- mirrors csrc/ude_tc_host.cu:tc_fill_layers."""
    r16 = lambda v: (v + 15) // 16 * 16
    fwd = dlt = dw = 0
    for sizes, head in ((spec.net_sizes, 2 * spec.n_regions), (spec.aug_sizes, 3 * spec.n_regions)):
        if not len(sizes):
            continue
        dims = [16] + [r16(v) for v in sizes] + [r16(head)]
        for li, (i, o) in enumerate(zip(dims[:-1], dims[1:])):
            if li < len(dims) - 2:          # the head layers' forward runs on the FMA pipe (fp32)
                fwd += 6 * 128 * i * o
            dlt += 6 * 128 * i * o
            dw += 8 * 16 * 64 * ((2 * i + 8) + (i + 8))
    return {"fwd": fwd, "bwd": fwd + dlt + dw}


def flops_per_traj_step(macs: int) -> dict:
    """Algorithmic flops: forward 2/MAC, backward 4/MAC (dX + dW), 4 stages;
    the adjoint's recomputation of the stage activations is NOT counted."""
    return {"fwd": 4 * 2 * macs, "bwd": 4 * 4 * macs, "total": 4 * 6 * macs}


# ---------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for p in self.samples:
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------
# CPU reference arm (oracle port; also the cpu_baseline leg of the GPU arm)
# ---------------------------------------------------------------------------------------
def cpu_reference_step(B: int, threads: int, reps: int):
    """fwd + backward of the reference path on the host: returns median seconds per pass."""
    import torch
    from oracle import odeint_ref
    from oracle.ude_rhs import OracleFunc, init_theta, synth_y0
    torch.set_num_threads(threads)
    spec = _spec()
    theta = init_theta(spec, seed=0)
    y0 = synth_y0(B, 1, 8, seed=0)
    t = torch.arange(float(N_STEPS_GRID + 1))
    week = torch.arange(N_STEPS_GRID + 1, dtype=torch.float32)
    y_ili = (0.02 + 0.5 * torch.exp(-((week - 5.0) / 4.0) ** 2)).reshape(-1, 1, 1)
    times = []
    for i in range(reps + 1):
        func = OracleFunc(spec, theta)
        y0g = y0.clone().requires_grad_(True)
        t0 = time.perf_counter()
        y = odeint_ref.odeint(func, y0g, t, method="rk4", options=dict(step_size=t[1] - t[0]))
        loss = _loss(y, y_ili, func)
        loss.backward()
        dt = time.perf_counter() - t0
        if i > 0:
            times.append(dt)
    return statistics.median(times)


def _loss(y, y_ili, func):
    """The synthetic loss of SURVEY.md 8d: ILI-shaped MSE on the I compartment, the reference's
    0.1 * latent_init_loss (lib/train_functions.py:116-126) on S,I,R, kl_p on posterior()
    (lib/VAE.py:172-177, lib/train_functions.py:77-80) and 0.1 * Fa_norm (lib/VAE.py:179-183)."""
    import torch
    from torch.distributions import Normal, kl_divergence
    x = y[..., :3]
    mse = torch.mean((y[..., 1] - y_ili.to(y.device)) ** 2)
    reg = 0.1 * (torch.where(x < 0, x.abs(), torch.zeros_like(x)) +
                 torch.where(x > 1, (1 - x).abs(), torch.zeros_like(x))).sum() / y.shape[1]
    post = func.posterior()
    prior = Normal(torch.tensor([0.8, 0.55], device=y.device), torch.tensor([0.2, 0.2], device=y.device))
    kl_p = kl_divergence(prior, post).mean()
    fa_norm = 0.1 * torch.norm(torch.stack(func.tracker)) / (y.shape[1] ** 0.5)
    return mse + reg + kl_p + fa_norm


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    secs = []
    for _ in range(args.warmup):
        cpu_reference_step(CPU_SAMPLE_B, cores, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        secs.append(cpu_reference_step(CPU_SAMPLE_B, cores, 1))
    sec = statistics.median(secs)
    value = CPU_SAMPLE_B * N_STEPS_GRID / sec
    sample = f"B={CPU_SAMPLE_B} trajectories x {N_STEPS_GRID} RK4 steps, fwd+autograd backward, {cores} torch threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[3]: FaFp R=1 L=8 [64,64,32]/[64,64], RK4 8 weekly steps, "
                               "fwd+adjoint; reference CPU path on a bounded sample", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import ude_b200
    from oracle.ude_rhs import init_theta, synth_y0
    from ude_b200 import _lib
    from ude_b200.odeint import _grid_on_device, _ptr, _stream_ptr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the fused UDE solve has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    L = ude_b200.lib()
    spec = _spec()
    B = args.batch
    theta_h = init_theta(spec, seed=0)
    mod = ude_b200.FaFp(1, 8, net_sizes=[64, 64, 32], aug_net_sizes=[64, 64])
    info = ude_b200.describe_func(mod)
    off = 0
    with torch.no_grad():
        for lin in info.linears():
            for p in (lin.weight, lin.bias):
                p.copy_(theta_h[off:off + p.numel()].reshape(p.shape))
                off += p.numel()
    mod = mod.to(dev)
    info = ude_b200.describe_func(mod)
    macs = int(L.ude_macs_per_eval(C.byref(info.desc(B))))
    fl = flops_per_traj_step(macs)

    # synthetic inputs: each rank owns its own shard of B trajectories (seeded by rank)
    y0_host = synth_y0(B, 1, 8, seed=100 + rank).pin_memory()
    y0 = y0_host.to(dev, non_blocking=True)
    t = torch.arange(float(N_STEPS_GRID + 1))
    week = torch.arange(N_STEPS_GRID + 1, dtype=torch.float32)
    y_ili = (0.02 + 0.5 * torch.exp(-((week - 5.0) / 4.0) ** 2)).reshape(-1, 1, 1).to(dev)
    (tables, n_steps) = _grid_on_device(t, t[1] - t[0], dev)
    n_out = t.numel()

    # ---- fixed cotangent for the device-resident leg: d(loss)/d(latent) of the synthetic loss
    y0g = y0.clone().requires_grad_(True)
    mod.clear_tracking()
    y = ude_b200.odeint(mod, y0g, t, method="rk4", options=dict(step_size=t[1] - t[0]))
    side_rows = mod.params[-1]
    loss = _loss(y, y_ili, mod)
    g_out = torch.autograd.grad(loss, y, retain_graph=False)[0].contiguous()
    g_side = torch.tensor([1e-8, -1e-8, 1e-9, 1e-9, 1e-9], dtype=torch.float32, device=dev)
    del y, loss, y0g, side_rows

    desc = info.desc(B)
    plan = _lib.UdePlan()
    _lib.check(L.ude_plan(C.byref(desc), 0, n_steps, C.byref(plan)), "ude_plan")
    theta = info.theta().detach().contiguous()
    y_out = torch.empty((n_out, B, 1, 8), dtype=torch.float32, device=dev)
    side = torch.empty(8, dtype=torch.float64, device=dev)
    ws = torch.empty(int(plan.workspace_bytes), dtype=torch.uint8, device=dev)
    tape = torch.empty(int(plan.tape_bytes), dtype=torch.uint8, device=dev)
    g_y0 = torch.empty_like(y0)
    g_theta = torch.empty_like(theta)
    dt_t, os_t, ok_t, ow_t = tables
    stream = _stream_ptr(dev)

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step_resident(timed: bool):
        """One pass with every input resident in HBM: fwd launch, bwd launch, grad all-reduce."""
        if timed:
            ev[0].record()
        st = L.ude_rk4_forward(C.byref(desc), _ptr(theta), _ptr(y0), n_steps, _ptr(dt_t), n_out, _ptr(os_t),
                               _ptr(ok_t), _ptr(ow_t), _ptr(y_out), _ptr(tape), _ptr(side), _ptr(ws),
                               int(plan.workspace_bytes), stream)
        _lib.check(st, "fwd")
        if timed:
            ev[1].record()
        st = L.ude_rk4_backward(C.byref(desc), _ptr(theta), _ptr(y0), n_steps, _ptr(dt_t), n_out, _ptr(os_t),
                                _ptr(ok_t), _ptr(ow_t), _ptr(tape), _ptr(g_out), _ptr(g_side), _ptr(g_y0),
                                _ptr(g_theta), _ptr(ws), int(plan.workspace_bytes), stream)
        _lib.check(st, "bwd")
        if timed:
            ev[2].record()
        if world > 1:
            dist.all_reduce(g_theta)
            dist.all_reduce(side)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step_resident(False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fwd_ms, bwd_ms = [], []
    e0.record()
    for _ in range(args.steps):
        step_resident(True)
        # per-kernel durations are read after the loop (events stay valid until re-recorded), so
        # only the last step's are kept when steps > 1; a short sync-free loop keeps the region tight.
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    fwd_ms = ev[0].elapsed_time(ev[1])
    bwd_ms = ev[1].elapsed_time(ev[2])
    clocks = sampler.stop()
    tm = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms = float(tm.item())
    ms_per_step = total_ms / args.steps
    value = world * B * n_steps / (ms_per_step * 1e-3)

    # ---- end-to-end leg: public API, host inputs, H2D + loss + backward + D2H every step
    def step_e2e():
        y0d = y0_host.to(dev, non_blocking=True).requires_grad_(True)
        mod.clear_tracking()
        mod.zero_grad(set_to_none=True)
        yy = ude_b200.odeint(mod, y0d, t, method="rk4", options=dict(step_size=t[1] - t[0]))
        ls = _loss(yy, y_ili, mod)
        ls.backward()
        flat = torch.cat([p.grad.reshape(-1) for p in mod.parameters()])
        if world > 1:
            dist.all_reduce(flat)
        out = torch.cat([ls.detach().reshape(1), flat]).to("cpu")
        return out

    e2e_steps = max(1, min(args.steps, 3))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = step_e2e()
    barrier()
    e2e_sec = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_sec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_sec = float(te.item())
    e2e_value = world * B * n_steps / e2e_sec

    if rank == 0:
        # ---- roofline of the dominant kernel (the adjoint): fp32 FMA pipe, measured peak
        peak = C.c_double(0.0)
        _lib.check(L.ude_measure_ffma_tflops(C.byref(peak), stream), "ffma peak")
        ach_bwd = B * n_steps * fl["bwd"] / (bwd_ms * 1e-3) / 1e12
        ach_fwd = B * n_steps * fl["fwd"] / (fwd_ms * 1e-3) / 1e12
        bytes_ts = 3 * 4 * 1 * 8          # fwd state store + bwd state load + bwd cotangent load
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        path = int(L.ude_rk4_path(C.byref(desc)))
        kname = "rk4_bwd_tc_kernel<1,8> (tcgen05)" if path == 1 else "rk4_bwd_kernel (FFMA)"
        tensor = None
        if path == 1:
            tm = tensor_macs_per_tile_eval(spec)
            evals = (B + 127) // 128 * n_steps * 4
            t_peak = peaks.get("bf16_tflops")
            t_bwd = 2.0 * tm["bwd"] * evals / (bwd_ms * 1e-3) / 1e12
            t_fwd = 2.0 * tm["fwd"] * evals / (fwd_ms * 1e-3) / 1e12
            tensor = {"note": "bf16 flops actually issued to the tensor pipe (6 bf16 products emulate one fp32 product; "
                              "M=128, N<=64 instructions sit on the 50-cycle issue floor measured by tools/umma_probe.cu)",
                      "bwd_issued_tflops": t_bwd, "fwd_issued_tflops": t_fwd, "peak": t_peak,
                      "peak_source": "MEASURED_PEAKS.json bf16_tflops (cuBLAS burst), of measured",
                      "bwd_frac": t_bwd / t_peak if t_peak else None, "fwd_frac": t_fwd / t_peak if t_peak else None}
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
                tr = json.load(f)
            if tr.get("batch") == B and tr.get("grid_steps") == n_steps:
                traffic = tr.get("bwd_dram_bytes_per_launch")
        except Exception:
            pass
        cpu = None
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sec = cpu_reference_step(CPU_SAMPLE_B, cores, 3)
            cpu = {"value": CPU_SAMPLE_B * N_STEPS_GRID / sec, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"B={CPU_SAMPLE_B} x {N_STEPS_GRID} steps fwd+autograd backward, median of 3, "
                             f"oracle port of lib/models.py + torchdiffeq rk4, {cores} torch threads"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[3]: synthetic ensemble, FaFp R=1 L=8 net [64,64,32] aug [64,64], "
                                   "RK4(3/8) t=[0..8] h=1 (8 steps), fused forward + discrete-adjoint backward"
                                   + (", NCCL all-reduce of dtheta" if world > 1 else ""),
                       "trajectories_per_gpu": B, "grid_steps": n_steps, "parallelism": f"trajectory-shard x{world}",
                       "l2": "inputs larger than L2 (y0 32 MiB + cotangent 288 MiB + tape 384 MiB per pass)",
                       "tile_traj": int(plan.tile_traj), "ctas": int(plan.n_ctas)},
            "roofline": {"bound": "fp32-fma", "kernel": kname, "achieved": ach_bwd, "peak": float(peak.value),
                         "unit": "TFLOP/s", "frac": ach_bwd / float(peak.value) if peak.value else None,
                         "peak_source": "measured live: ude_measure_ffma_tflops (register-resident FFMA chains)",
                         "traffic": traffic,
                         "algorithmic_bytes_per_launch": B * n_steps * (48 + 32) + B * 64,
                         "tensor_pipe": tensor,
                         "fwd_kernel": {"achieved": ach_fwd, "frac": ach_fwd / float(peak.value) if peak.value else None,
                                        "ms": fwd_ms},
                         "bwd_ms": bwd_ms,
                         "whole_step_frac": (B * n_steps * fl["total"] / (ms_per_step * 1e-3) / 1e12) / float(peak.value)
                         if peak.value else None,
                         "hbm_gbs_algorithmic": B * n_steps * bytes_ts / (ms_per_step * 1e-3) / 1e9,
                         "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(y0_host.numel() * 4),
                    "d2h_bytes_per_step": int(res.numel() * 4), "ms_per_step": e2e_sec * 1e3},
            "gpu_launches": 6 * args.steps,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="trajectories per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
