#!/usr/bin/env python
"""bench.py — UDE trajectory-steps/s (forward + discrete adjoint) on N B200s.

Contract (driver): ``python bench.py --gpus N --steps K --warmup W`` (torchrun for N>1)
prints ONE JSON line on rank 0.  A "step" is one pass of the hot path over one batch of
synthetic ILI-shaped input: the fused RK4 forward over the whole time grid followed by the
fused discrete-adjoint backward (+ the NCCL all-reduces of the side sums and of dL/dtheta when N>1).

Workload = BASELINE.json configs[3] (the configuration the metric is quoted on):
  FaFp(R=1, L=8, net [64,64,32], aug [64,64]), B = 2**20 trajectories IN TOTAL, sharded by trajectory
  over the N ranks (strong scaling, as the metric says: "B=1M ... trajectory-sharded at 2/4/8 GPUs");
  t = [0..8] weeks, h = 1 (8 RK4 steps).  The weak-scaling figure (2**20 per GPU) rides along as a sub-record.

``secondary`` (N = 1 only) carries the other BASELINE configs: configs[0] single-trajectory latency, the
reference's own training shape (B = 2048), configs[2] (49 states x 256 samples, 28 daily steps), configs[1]
(49-state Dopri5 epoch) and configs[4] (64 weight-set groups per GPU).

``--impl reference`` times the reference's CPU path for the same metric/config on the host cores: the real
``lib.models`` classes when /root/reference is importable (kind "reference"), else the oracle port of them
(kind "port"), driven by the oracle's restatement of torchdiffeq's rk4 (torchdiffeq itself is not installable
offline) with torch autograd, on a bounded sample.
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
B_TOTAL = 1 << 20
CPU_SAMPLE_B = 65536
METRIC = "ude_trajectory_steps_per_sec_fwd_plus_adjoint"
UNIT = "trajectory-steps/s"
WORKLOAD = ("BASELINE configs[3]: synthetic ensemble, FaFp R=1 L=8 net [64,64,32] aug [64,64] (nn.Linear default init, "
            "aug head scaled by aug_head_scale=0.05 so that trajectories stay inside the RHS clamp), RK4(3/8) t=[0..8] h=1 "
            "(8 steps), fused forward + discrete-adjoint backward; loss = MSE of the Decoder Linear(3->1) output against an ILI-shaped season + "
            "0.1 latent_init_loss + kl_p(posterior) + 0.1 Fa_norm (lib/VAE.py:138-186)")


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


def fma_pipe_macs_per_traj_eval(spec) -> dict:
    """MACs of the tcgen05 kernels that stay on the fp32 FMA pipe (csrc/ude_tc.cuh): the head layers forward, the heads'
    Wh^T g and the first layers' W0^T delta0 in the sweep."""
    heads = first = 0
    for sizes, head in ((spec.net_sizes, 2 * spec.n_regions), (spec.aug_sizes, 3 * spec.n_regions)):
        if len(sizes):
            heads += sizes[-1] * head
            first += spec.n_regions * spec.latent_dim * sizes[0]
    return {"fwd": heads, "bwd": heads + heads + first}


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
# CPU reference arm (also the cpu_baseline leg of the GPU arm)
# ---------------------------------------------------------------------------------------
def reference_func(spec, theta, fa_w=1.0):
    """(callable ODE func, kind, theta-ordered parameter list): the REAL reference class
    (/root/reference/lib/models.py:109-265) holding `theta` when the reference tree is importable, else the oracle's
    port of it.  Both follow the ODE-func protocol (params / tracker / posterior())."""
    import torch
    ref_root = "/root/reference"
    if os.path.isdir(os.path.join(ref_root, "lib")):
        try:
            if ref_root not in sys.path:
                sys.path.insert(0, ref_root)
            import lib.models as ref_models
            kw = {}
            if spec.has_rate:
                kw["net_sizes"] = list(spec.net_sizes)
            if spec.has_aug:
                kw["aug_net_sizes"] = list(spec.aug_sizes)
            mod = getattr(ref_models, spec.kind)(spec.n_regions, spec.latent_dim, **kw)
            nets = []
            if spec.kind == "Fp":
                nets = [mod.Fp_net]
            elif spec.kind == "FaFp":
                nets = [mod.net, mod.aug_net]
            else:
                nets = [mod.aug_net]
            off = 0
            plist = []
            with torch.no_grad():
                for net in nets:
                    for m in net:
                        if isinstance(m, torch.nn.Linear):
                            for p in (m.weight, m.bias):
                                p.copy_(theta[off:off + p.numel()].reshape(p.shape))
                                off += p.numel()
                                plist.append(p)
            assert off == theta.numel()
            if spec.kind == "FaFp":
                mod.Fa_w = fa_w
            return mod, "reference", plist
        except Exception:
            pass
    from oracle.ude_rhs import OracleFunc
    f = OracleFunc(spec, theta, fa_w)
    return f, "port", [f.theta]


def cpu_reference_step(B: int, threads: int, reps: int, spec=None, t=None, step_size=None):
    """fwd + backward of the reference path on the host: returns (median seconds per pass, kind)."""
    import torch
    from oracle import odeint_ref
    from oracle.ude_rhs import init_theta, synth_y0
    torch.set_num_threads(threads)
    spec = spec or _spec()
    theta = init_theta(spec, seed=0)
    y0 = synth_y0(B, spec.n_regions, spec.latent_dim, seed=0)
    if t is None:
        t = torch.arange(float(N_STEPS_GRID + 1))
    h = (t[1] - t[0]) if step_size is None else step_size
    week = t.clone()
    y_ili = (0.02 + 0.5 * torch.exp(-((week - 5.0) / 4.0) ** 2)).reshape(-1, 1, 1)
    times, kind = [], "port"
    dec = _decoder("cpu") if spec.n_regions == 1 else None       # the Decoder of lib/models.py:26-51 (torch ops on the trajectories)
    for i in range(reps + 1):
        if dec is not None:
            dec.zero_grad(set_to_none=True)
        func, kind, _ = reference_func(spec, theta)
        y0g = y0.clone().requires_grad_(True)
        t0 = time.perf_counter()
        y = odeint_ref.odeint(func, y0g, t, method="rk4", options=dict(step_size=h))
        loss = _loss(y, y_ili, func, dec)
        loss.backward()
        dt = time.perf_counter() - t0
        if i > 0:
            times.append(dt)
    return statistics.median(times), kind


DEC_W, DEC_B = [[0.05, 1.0, -0.03]], [0.01]     # the headline's Decoder Linear(3 -> 1) (lib/models.py:26-51), fixed values


def _decoder(dev):
    import torch
    lin = torch.nn.Linear(3, 1)
    with torch.no_grad():
        lin.weight.copy_(torch.tensor(DEC_W))
        lin.bias.copy_(torch.tensor(DEC_B))
    return lin.to(dev)


def _loss(y, y_ili, func, dec=None):
    """The synthetic loss of SURVEY.md 8d: ILI-shaped MSE on the decoded prediction (lib/VAE.py:138,161; on the I
    compartment when no decoder is given), the reference's 0.1 * latent_init_loss (lib/train_functions.py:116-126) on
    S,I,R, kl_p on posterior() (lib/VAE.py:172-177, lib/train_functions.py:77-80) and 0.1 * Fa_norm (lib/VAE.py:179-183)."""
    import torch
    x = y[..., :3]
    pred = y[..., 1] if dec is None else dec(x.reshape(y.shape[0], y.shape[1], -1))
    y_ili = y_ili.to(y.device)
    reg = (torch.where(x < 0, x.abs(), torch.zeros_like(x)) + torch.where(x > 1, (1 - x).abs(), torch.zeros_like(x))).sum()
    return _loss_tail(pred, y_ili if dec is None else y_ili.reshape(-1, 1, 1), reg, y.shape[1], func)


_PRIORS = {}


def _prior(device):
    """N([0.8, 0.55], [0.2, 0.2]) (lib/VAE.py:42-43), built once per device: creating tensors from python lists is a host -> device
    copy, which a CUDA-graph capture of the iteration must not contain."""
    import torch
    from torch.distributions import Normal
    key = str(device)
    if key not in _PRIORS:
        _PRIORS[key] = Normal(torch.tensor([0.8, 0.55], device=device), torch.tensor([0.2, 0.2], device=device), validate_args=False)
    return _PRIORS[key]


def _loss_tail(pred, y_ili, reg_sum, B, func):
    """mse(pred, y_ili) + 0.1 * reg_sum / B + kl_p + 0.1 * Fa_norm / sqrt(B): shared by the plain path (pred / reg from torch ops on
    the trajectories) and the fused-consumer path (pred / reg straight from ude_b200.odeint_decoded)."""
    import torch
    from torch.distributions import Normal, kl_divergence
    y = pred
    mse = torch.mean((pred - y_ili) ** 2)
    loss = mse + 0.1 * reg_sum / B
    if getattr(func, "params", None):
        post = func.posterior()
        loss = loss + kl_divergence(_prior(pred.device), post).mean()
    if getattr(func, "tracker", None):
        loss = loss + 0.1 * torch.norm(torch.stack(func.tracker)) / (B ** 0.5)
    return loss


def cpu_baseline_record(threads_all: int):
    """The reference path on the host at all threads and at the reference's own torch.set_num_threads(1)
    (run_ode.py:28), on a bounded B = 65 536 sample of the bench workload."""
    sec_all, kind = cpu_reference_step(CPU_SAMPLE_B, threads_all, 3)
    sec_one, _ = cpu_reference_step(CPU_SAMPLE_B, 1, 1)
    what = ("the real lib.models.FaFp" if kind == "reference" else "oracle port of lib/models.py (reference tree absent on this box)")
    return {"value": CPU_SAMPLE_B * N_STEPS_GRID / sec_all, "unit": UNIT, "cores": threads_all, "kind": kind,
            "sample": f"B={CPU_SAMPLE_B} x {N_STEPS_GRID} steps fwd+autograd backward, median of 3; {what} driven by the "
                      f"oracle's torchdiffeq-rk4 restatement, {threads_all} torch threads",
            "single_thread": {"value": CPU_SAMPLE_B * N_STEPS_GRID / sec_one, "cores": 1,
                              "note": "torch.set_num_threads(1) as the reference sets it (run_ode.py:28); one pass"}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    secs, kind = [], "port"
    for _ in range(args.warmup):
        cpu_reference_step(CPU_SAMPLE_B, cores, 1)
    for _ in range(args.steps):
        s, kind = cpu_reference_step(CPU_SAMPLE_B, cores, 1)
        secs.append(s)
    sec = statistics.median(secs)
    value = CPU_SAMPLE_B * N_STEPS_GRID / sec
    sec_one, _ = cpu_reference_step(CPU_SAMPLE_B, 1, 1)
    what = "the real lib.models.FaFp" if kind == "reference" else "oracle port of lib/models.py (reference tree absent on this box)"
    sample = (f"B={CPU_SAMPLE_B} trajectories x {N_STEPS_GRID} RK4 steps, fwd+autograd backward, {what}, {cores} torch threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + "; reference CPU path on a bounded sample", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "single_thread": {"value": CPU_SAMPLE_B * N_STEPS_GRID / sec_one, "cores": 1,
                                           "note": "torch.set_num_threads(1), run_ode.py:28"}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def _cuda_timed(fn, reps=3, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _wall_timed(fn, reps=5, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def _module(spec, theta, dev, fa_w=1.0):
    import torch

    import ude_b200
    kw = {}
    if spec.has_rate:
        kw["net_sizes"] = list(spec.net_sizes)
    if spec.has_aug:
        kw["aug_net_sizes"] = list(spec.aug_sizes)
    mod = getattr(ude_b200, spec.kind)(spec.n_regions, spec.latent_dim, **kw)
    info = ude_b200.describe_func(mod)
    off = 0
    with torch.no_grad():
        for lin in info.linears():
            for p in (lin.weight, lin.bias):
                p.copy_(theta[off:off + p.numel()].reshape(p.shape))
                off += p.numel()
    return mod.to(dev)


def secondary_configs(dev, ffma_peak_tflops: float, with_cpu: bool) -> dict:
    """The other BASELINE.json configs, through the public API (odeint + loss.backward()), 1 GPU.  Times are per
    ITERATION (fused forward + fused adjoint + the autograd glue); `frac_ffma` = algorithmic flops (SURVEY 8d:
    24 x MACs per trajectory-step, first layers at their full R*L width) / time / measured FFMA peak."""
    import torch

    import ude_b200
    from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
    out = {}
    peak = ffma_peak_tflops * 1e12

    def fwd_bwd(mod, y0, t, **kw):
        def step():
            y0g = y0.clone().requires_grad_(True)
            mod.clear_tracking()
            mod.zero_grad(set_to_none=True)
            y = ude_b200.odeint(mod, y0g, t, **kw)
            (y[..., 1] ** 2).mean().backward()
        return step

    # ---- configs[0]: one trajectory, RK4 fwd + backprop: launch-latency bound by construction
    spec1 = _spec()
    mod1 = _module(spec1, init_theta(spec1, 0), dev)
    lat = {}
    for name, t in (("weekly_8_steps", torch.arange(9.)), ("daily_85_steps", torch.arange(86.) / 7)):
        y0 = synth_y0(1, 1, 8, seed=5).to(dev)
        n = t.numel() - 1
        ms = _wall_timed(fwd_bwd(mod1, y0, t, method="rk4", options=dict(step_size=t[1] - t[0])), reps=10, warm=3)
        rec = {"ms_per_solve_fwd_bwd": ms, "us_per_step": ms * 1e3 / n, "grid_steps": n}
        if with_cpu:
            sec, kind = cpu_reference_step(1, 1, 2, spec1, t)
            rec["cpu_ms_per_solve_1_thread"] = sec * 1e3
            rec["cpu_kind"] = kind
        lat[name] = rec
    out["configs[0] B=1 FaFp R=1 RK4 fwd+backprop (run_ode.py:93,130,147), host wall clock per call"] = lat

    # ---- the reference's own training shape: 32 x 64 = 2048 trajectories, 7 weekly steps (run_ode.ipynb:206-213)
    y0 = synth_y0(2048, 1, 8, seed=6).to(dev)
    t = torch.arange(8.)
    step = fwd_bwd(mod1, y0, t, method="rk4", options=dict(step_size=t[1] - t[0]))
    rec = {"ms_per_iter_wall": _wall_timed(step, reps=10, warm=3), "ms_per_iter_device": _cuda_timed(step, reps=10, warm=3)}
    rec["trajectory_steps_per_s"] = 2048 * 7 / (rec["ms_per_iter_wall"] * 1e-3)
    if with_cpu:
        sec, kind = cpu_reference_step(2048, 1, 2, spec1, t)
        rec["cpu_ms_per_iter_1_thread"] = sec * 1e3
        rec["cpu_kind"] = kind
    # the same iteration + a capturable Adam step replayed from ONE CUDA graph (ude_b200.GraphedIteration); a fresh
    # module: its gradient buffers are created on the capture stream
    try:
        mod_g = _module(spec1, init_theta(spec1, 0), dev)
        opt = torch.optim.Adam(mod_g.parameters(), lr=1e-3, capturable=True)

        def step_fn(y):
            mod_g.clear_tracking()
            yy = ude_b200.odeint(mod_g, y, t, method="rk4", options=dict(step_size=t[1] - t[0]))
            return (yy[..., 1] ** 2).mean()
        it = ude_b200.GraphedIteration(step_fn, list(mod_g.parameters()), opt, grad_lim=300.0)
        rec["ms_per_iter_cuda_graph_incl_adam"] = _cuda_timed(lambda: it(y0), reps=20, warm=3)
        rec["ms_per_iter_cuda_graph_wall"] = _wall_timed(lambda: it(y0), reps=20, warm=3)
    except Exception as e:
        rec["cuda_graph_error"] = f"{type(e).__name__}: {e}"
    out["training shape B=2048 FaFp R=1, 7 weekly RK4 steps, fwd+backprop"] = rec

    # ---- configs[3] variants of SURVEY 8d: the headline model on the 56-step daily grid, and the rate-only Fp([64, 64]) model
    B3 = 1 << 20
    y0 = synth_y0(B3, 1, 8, seed=7).to(dev)
    t = torch.arange(57.) / 7
    ms = _cuda_timed(fwd_bwd(mod1, y0, t, method="rk4", options=dict(step_size=t[1] - t[0])), reps=2, warm=1)
    ts = B3 * 56 / (ms * 1e-3)
    out["configs[3] variant: FaFp R=1, B=2^20, 56 daily RK4 steps (t = arange(57)/7, h = t[1]-t[0]) fwd+adjoint"] = {
        "ms_per_iter": ms, "trajectory_steps_per_s": ts, "frac_ffma": ts * 24 * spec1.macs_per_eval() / peak}
    specp = UdeSpec("Fp", 1, 8, (64, 64), ())
    modp = _module(specp, init_theta(specp, 0), dev)
    t = torch.arange(9.)
    ms = _cuda_timed(fwd_bwd(modp, y0, t, method="rk4", options=dict(step_size=1.0)))
    ts = B3 * 8 / (ms * 1e-3)
    out["configs[3] variant: Fp([64, 64]) R=1, B=2^20, 8 weekly RK4 steps fwd+adjoint"] = {
        "ms_per_iter": ms, "trajectory_steps_per_s": ts, "frac_ffma": ts * 24 * specp.macs_per_eval() / peak}
    del y0, modp

    # ---- configs[2]: 49 states x 256 latent samples (x b), 28 daily RK4 steps
    spec49 = UdeSpec("FaFp", 49, 8, (64, 64, 32), (64, 64))
    mod49 = _module(spec49, init_theta(spec49, 0), dev)
    t = torch.arange(29.) / 7
    for b in (1, 32):
        B = 256 * b
        y0 = synth_y0(B, 49, 8, seed=1).to(dev)
        ms = _cuda_timed(fwd_bwd(mod49, y0, t, method="rk4", options=dict(step_size=t[1] - t[0])))
        ts = B * 28 / (ms * 1e-3)
        out[f"configs[2] FaFp R=49 y0 [{B},49,8] RK4 28 daily steps fwd+adjoint"] = {
            "ms_per_iter": ms, "trajectory_steps_per_s": ts, "frac_ffma": ts * 24 * spec49.macs_per_eval() / peak}

    # ---- configs[1]: 49-state regional batch, Dopri5 rtol 1e-5, one epoch = 153 iterations of [2048, 49, 8]
    y0 = synth_y0(2048, 49, 8, seed=2).to(dev)
    t = torch.arange(9.)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ms = _cuda_timed(fwd_bwd(mod49, y0, t, rtol=1e-5, atol=1e-9, method="dopri5", options=dict(per_trajectory=True)))
        with torch.no_grad():
            _, st = ude_b200.odeint(mod49, y0, t, rtol=1e-5, atol=1e-9, method="dopri5",
                                    options=dict(per_trajectory=True, return_stats=True))
    n_acc, n_att = float(st.n_accept.float().sum()), float(st.n_attempts.float().sum())
    # forward: 6 evaluations per attempt (+ f0 + the initial-step probe) at 2 flop/MAC; adjoint: 6 per ACCEPTED step (+ f0) at 4
    flops = (2 * (6 * n_att + 2 * 2048) + 4 * (6 * n_acc + 2048)) * spec49.macs_per_eval()
    rec1 = {
        "ms_per_iter": ms, "epoch_153_iters_s": ms * 153e-3, "accepted_steps_mean": n_acc / 2048, "attempted_steps_mean": n_att / 2048,
        "accepted_trajectory_steps_per_s": n_acc / (ms * 1e-3), "frac_ffma": flops / (ms * 1e-3) / peak,
        "controller": "one per trajectory (options['per_trajectory'] = True)"}
    # torchdiffeq's own batch semantics: ONE error norm / step size / decision for all 2048 x 49 x 8 states (the default mode at this size)
    ms_s = _cuda_timed(fwd_bwd(mod49, y0, t, rtol=1e-5, atol=1e-9, method="dopri5", options=dict(per_trajectory=False)))
    with torch.no_grad():
        _, st_s = ude_b200.odeint(mod49, y0, t, rtol=1e-5, atol=1e-9, method="dopri5", options=dict(per_trajectory=False, return_stats=True))
    rec1["shared_step"] = {"ms_per_iter": ms_s, "epoch_153_iters_s": ms_s * 153e-3, "accepted_steps": int(st_s.n_accept[0]),
                           "attempted_steps": int(st_s.n_attempts[0]),
                           "controller": "torchdiffeq batch semantics: one RMS norm over the whole tensor, cooperative launch"}
    out["configs[1] FaFp R=49 y0 [2048,49,8] Dopri5 rtol 1e-5 atol 1e-9 t=[0..8] fwd+adjoint"] = rec1

    # ---- configs[4]: 64 weight-set groups per GPU (512 over 8), Fp hidden 32, 2048 trajectories each, 56 outputs
    specg = UdeSpec("Fp", 1, 8, (32, 32), ())
    G, Bg = 64, 2048
    mods = [_module(specg, init_theta(specg, g), dev) for g in range(G)]
    y0 = synth_y0(G * Bg, 1, 8, seed=3).reshape(G, Bg, 1, 8).to(dev)
    t = torch.linspace(1, 56, 56) / 7

    def step_g():       # 64 separate modules: their 6 x 64 parameter tensors are gathered / scattered by torch every iteration
        y0g = y0.clone().requires_grad_(True)
        for m in mods:
            m.zero_grad(set_to_none=True)
        y = ude_b200.odeint_groups(mods, y0g, t, method="rk4", options=dict(step_size=1.0))
        (y[..., 1] ** 2).mean().backward()

    ms_sep = _cuda_timed(step_g)
    flat = ude_b200.flatten_group_parameters(mods)       # the same weights as ONE [G, n_theta] leaf: no gather, one gradient tensor

    def step_flat():
        y0g = y0.clone().requires_grad_(True)
        flat.grad = None
        y = ude_b200.odeint_groups(mods, y0g, t, method="rk4", options=dict(step_size=1.0), theta=flat)
        (y[..., 1] ** 2).mean().backward()

    ms = _cuda_timed(step_flat)
    ts = G * Bg * 8 / (ms * 1e-3)
    out["configs[4] 64 groups x Fp(R=1,L=8,[32,32]) x 2048 trajectories, RK4 h=1 t=[1/7..8] (8 grid steps, 56 outputs) fwd+adjoint"] = {
        "ms_per_iter": ms, "trajectory_steps_per_s": ts, "frac_ffma": ts * 24 * specg.macs_per_eval() / peak,
        "how": "one forward + one adjoint launch over the tiles of all groups (tcgen05 engine), weights as one [G, n_theta] leaf "
               "(ude_b200.flatten_group_parameters)",
        "ms_per_iter_separate_module_parameters": ms_sep}
    return out


def dist_check(mod, dev, world, rank) -> str:
    """Untimed: a 4096-trajectory batch solved SHARDED (side sums all-reduced through options['process_group'],
    gradients through dist.all_reduce_grads) must give every rank the single-GPU gradient."""
    import torch
    import torch.distributed as dist

    import ude_b200
    from oracle.ude_rhs import synth_y0
    n_samples, batch = 64, 64
    y0 = synth_y0(n_samples * batch, 1, 8, seed=77).to(dev)
    t = torch.arange(5.)
    y_ili = (0.02 + 0.5 * torch.exp(-((t - 2.0) / 4.0) ** 2)).reshape(-1, 1, 1).to(dev)
    # single-GPU result (every rank computes it: same inputs, same kernels)
    mod.zero_grad(set_to_none=True)
    mod.clear_tracking()
    y = ude_b200.odeint(mod, y0.clone().requires_grad_(True), t, method="rk4", options=dict(step_size=1.0, kernels="ffma"))
    (_loss(y, y_ili, mod) * 1.0).backward()
    g_single = torch.cat([p.grad.reshape(-1) for p in mod.parameters()]).clone()
    # sharded: per-rank mean terms are rescaled to the global batch, side-channel terms see the all-reduced sums
    mod.zero_grad(set_to_none=True)
    mod.clear_tracking()
    y0s = ude_b200.dist.shard_flat(y0, n_samples, batch, world, rank).clone().requires_grad_(True)
    ys = ude_b200.odeint(mod, y0s, t, method="rk4", options=dict(step_size=1.0, kernels="ffma", process_group=True))
    x = ys[..., :3]
    local = (torch.sum((ys[..., 1] - y_ili) ** 2) / (y.shape[1] * y.shape[0] * y.shape[2]) +
             0.1 * (torch.where(x < 0, x.abs(), torch.zeros_like(x)) +
                    torch.where(x > 1, (1 - x).abs(), torch.zeros_like(x))).sum() / y.shape[1])
    from torch.distributions import Normal, kl_divergence
    prior = Normal(torch.tensor([0.8, 0.55], device=dev), torch.tensor([0.2, 0.2], device=dev))
    shared = kl_divergence(prior, mod.posterior()).mean() + 0.1 * torch.norm(torch.stack(mod.tracker)) / (y.shape[1] ** 0.5)
    # `shared` is the same global scalar on every rank: its parameter gradients are per-rank PARTIAL sums already
    # (all_reduce_side's backward is the identity), so it enters once per rank and the all-reduce completes it
    (local + shared).backward()
    ude_b200.dist.all_reduce_grads(mod.parameters())
    g_shard = torch.cat([p.grad.reshape(-1) for p in mod.parameters()])
    err = float((g_shard - g_single).abs().max() / g_single.abs().max())
    flag = torch.tensor([err], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    err = float(flag)
    return f"ok (max relative gradient difference {err:.1e})" if err <= 1e-5 else f"FAILED (max relative gradient difference {err:.3e})"


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
    mod = _module(spec, init_theta(spec, seed=0), dev)
    info = ude_b200.describe_func(mod)
    t = torch.arange(float(N_STEPS_GRID + 1))
    y_ili = (0.02 + 0.5 * torch.exp(-((t - 5.0) / 4.0) ** 2)).reshape(-1, 1, 1).to(dev)
    dec = _decoder(dev)
    (tables, n_steps) = _grid_on_device(t, t[1] - t[0], dev)
    n_out = t.numel()
    dt_t, os_t, ok_t, ow_t = tables
    stream = _stream_ptr(dev)

    def measure(B: int, steps: int, warmup: int, seed: int):
        """Device-resident leg at B trajectories on this rank: returns (ms_per_step max over ranks, fwd ms, bwd ms, plan)."""
        y0_host = synth_y0(B, 1, 8, seed=seed).pin_memory()
        y0 = y0_host.to(dev, non_blocking=True)
        # fixed cotangent: d(loss)/d(latent) of the synthetic loss
        y0g = y0.clone().requires_grad_(True)
        mod.clear_tracking()
        y = ude_b200.odeint(mod, y0g, t, method="rk4", options=dict(step_size=t[1] - t[0]))
        loss = _loss(y, y_ili, mod, dec)
        g_out = torch.autograd.grad(loss, y, retain_graph=False)[0].contiguous()
        g_side = torch.tensor([1e-8, -1e-8, 1e-9, 1e-9, 1e-9], dtype=torch.float32, device=dev)
        del y, loss, y0g
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
        n_ev = steps
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_ev)]

        def step(ev=None):
            if ev:
                ev[0].record()
            st = L.ude_rk4_forward(C.byref(desc), _ptr(theta), _ptr(y0), n_steps, _ptr(dt_t), n_out, _ptr(os_t),
                                   _ptr(ok_t), _ptr(ow_t), _ptr(y_out), _ptr(tape), _ptr(side), _ptr(ws),
                                   int(plan.workspace_bytes), stream)
            _lib.check(st, "fwd")
            if ev:
                ev[1].record()
            if world > 1:       # the loss needs the GLOBAL side sums before backward; in-stream: an NCCL kernel that overlaps the
                dist.all_reduce(side)   # persistent adjoint kernel takes an SM away from its one-CTA-per-SM grid
            st = L.ude_rk4_backward(C.byref(desc), _ptr(theta), _ptr(y0), n_steps, _ptr(dt_t), n_out, _ptr(os_t),
                                    _ptr(ok_t), _ptr(ow_t), _ptr(tape), _ptr(g_out), _ptr(g_side), _ptr(g_y0),
                                    _ptr(g_theta), _ptr(ws), int(plan.workspace_bytes), stream)
            _lib.check(st, "bwd")
            if ev:
                ev[2].record()
            if world > 1:
                dist.all_reduce(g_theta)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)

        for _ in range(warmup):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(evs[i])
        e1.record()
        barrier()
        total_ms = e0.elapsed_time(e1)
        fwd_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
        bwd_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
        tm = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return float(tm.item()) / steps, fwd_ms, bwd_ms, plan, desc, y0_host

    warmup = max(args.warmup, 3)
    B_rank = args.batch if args.batch else B_TOTAL // world
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_per_step, fwd_ms, bwd_ms, plan, desc, y0_host = measure(B_rank, args.steps, warmup, 100 + rank)
    clocks = sampler.stop()
    value = world * B_rank * n_steps / (ms_per_step * 1e-3)
    B = B_rank
    macs = int(L.ude_macs_per_eval(C.byref(desc)))
    fl = flops_per_traj_step(macs)

    weak = None
    if world > 1 and not args.batch:
        w_ms, _, _, _, _, _ = measure(B_TOTAL, min(args.steps, 3), 3, 200 + rank)
        weak = {"trajectories_per_gpu": B_TOTAL, "ms_per_step": w_ms, "value": world * B_TOTAL * n_steps / (w_ms * 1e-3),
                "unit": UNIT, "scaling": "weak"}
    check = dist_check(mod, dev, world, rank) if world > 1 else None

    # ---- end-to-end legs: public API, host inputs, H2D + loss + backward + D2H every step.
    #   fused consumers (the headline e2e): ude_b200.odeint_decoded -- decoder + latent_init_loss inside the solve
    #   plain: ude_b200.odeint, then the same decoder / loss as torch ops on the [T,B,R,L] trajectories (what lib/VAE.py does)
    def step_e2e(fused: bool):
        y0d = y0_host.to(dev, non_blocking=True).requires_grad_(True)
        mod.clear_tracking()
        mod.zero_grad(set_to_none=True)
        dec.zero_grad(set_to_none=True)
        opts = dict(step_size=t[1] - t[0])
        if world > 1:
            opts["process_group"] = True
        if fused:
            yp, reg = ude_b200.odeint_decoded(mod, dec, y0d, t, options=opts)
            ls = _loss_tail(yp, y_ili, reg, yp.shape[1], mod)
        else:
            yy = ude_b200.odeint(mod, y0d, t, method="rk4", options=opts)
            ls = _loss(yy, y_ili, mod, dec)
        ls.backward()
        flat = torch.cat([p.grad.reshape(-1) for p in (*mod.parameters(), *dec.parameters())])
        if world > 1:
            dist.all_reduce(flat)
        out = torch.cat([ls.detach().reshape(1), flat]).to("cpu")
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def time_e2e(fused: bool):
        e2e_steps = max(1, min(args.steps, 3))
        step_e2e(fused)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            r = step_e2e(fused)
        barrier()
        sec = (time.perf_counter() - t0) / e2e_steps
        te = torch.tensor([sec], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item()), r

    plain_sec, res_plain = time_e2e(False)
    fused_sec, res = time_e2e(True)
    e2e_agree = float((res - res_plain).abs().max() / res_plain.abs().max())      # both legs compute the same loss and gradients
    e2e_api = ("ude_b200.odeint_decoded (solve + Decoder + latent_init_loss fused, SURVEY 8f2), loss tail in torch, backward, "
               "gradients to host")
    e2e_sec = fused_sec
    graphed = None
    if not args.no_graph_e2e:
        # the same iteration through the package's iteration driver: H2D copy of y0 from pinned memory, ONE CUDA-graph
        # replay (fused solve, loss tail, adjoint, gradient all-reduce), D2H read of [loss | gradients].  Fresh modules: their
        # gradient accumulators must be created on the capture's stream.
        try:
            mod_g = _module(spec, init_theta(spec, seed=0), dev)
            dec_g = _decoder(dev)
            params_g = [*mod_g.parameters(), *dec_g.parameters()]
            opts_g = dict(step_size=t[1] - t[0])
            if world > 1:
                opts_g["process_group"] = True

            def step_fn(y0d):
                mod_g.clear_tracking()
                yp, reg = ude_b200.odeint_decoded(mod_g, dec_g, y0d, t, options=opts_g)
                return _loss_tail(yp, y_ili, reg, yp.shape[1], mod_g)

            def after(loss):
                flat = torch.cat([p.grad.reshape(-1) for p in params_g])
                if world > 1:
                    dist.all_reduce(flat)
                return torch.cat([loss.reshape(1), flat])

            it = ude_b200.GraphedIteration(step_fn, params_g, after_backward=after, device=dev)
            y0_req = y0_host.requires_grad_(True)

            def step_graph():
                it(y0_req)
                return it.last_extra.to("cpu")

            step_graph()
            barrier()
            g_steps = max(1, min(args.steps, 5))
            t0 = time.perf_counter()
            for _ in range(g_steps):
                res_g = step_graph()
            barrier()
            sec = (time.perf_counter() - t0) / g_steps
            tg = torch.tensor([sec], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            graphed = {"sec": float(tg.item()), "max_rel_diff_vs_eager": float((res_g - res).abs().max() / res.abs().max())}
            if graphed["max_rel_diff_vs_eager"] <= 1e-4:
                e2e_sec, res = graphed["sec"], res_g
                e2e_api = ("ude_b200.GraphedIteration (SURVEY 8f4) over ude_b200.odeint_decoded (8f2): pinned-host y0 -> H2D -> one CUDA-graph "
                           "replay (fused solve + Decoder + latent_init_loss, loss tail, adjoint" + (", NCCL all-reduces" if world > 1 else "") +
                           ") -> D2H of [loss | gradients]")
        except Exception as e:
            graphed = {"error": f"{type(e).__name__}: {e}"[:300]}
    e2e_value = world * B * n_steps / e2e_sec

    if rank == 0:
        peak = C.c_double(0.0)
        _lib.check(L.ude_measure_ffma_tflops(C.byref(peak), stream), "ffma peak")
        ffma_peak = float(peak.value)
        ach_bwd = B * n_steps * fl["bwd"] / (bwd_ms * 1e-3) / 1e12
        ach_fwd = B * n_steps * fl["fwd"] / (fwd_ms * 1e-3) / 1e12
        bytes_ts = 3 * 4 * 1 * 8          # fwd state store + bwd state load + bwd cotangent load
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        t_peak = peaks.get("bf16_tflops") or 1590.0
        t_peak_src = "MEASURED_PEAKS.json bf16_tflops (cuBLAS burst), of measured" if peaks.get("bf16_tflops") else "fallback 1590 (B200_PROFILING.md), of fallback"
        path = int(L.ude_rk4_path(C.byref(desc)))
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
                tr = json.load(f)
            if tr.get("batch") == B and tr.get("grid_steps") == n_steps:
                traffic = tr.get("bwd_dram_bytes_per_launch")
        except Exception:
            pass
        if path == 1:
            # Mixed bound of the tcgen05 adjoint: the time the two pipes it actually uses would need at their
            # measured peaks -- bf16 flops ISSUED to the tensor pipe / cuBLAS bf16 peak, and the fp32 FMA-pipe remainder (heads,
            # first-layer products) / measured FFMA peak -- the larger of the two is the bound; frac = bound / measured time.
            tm_ = tensor_macs_per_tile_eval(spec)
            fm_ = fma_pipe_macs_per_traj_eval(spec)
            evals = (B + 127) // 128 * n_steps * 4
            t_tensor = {k: 2.0 * tm_[k] * evals / (t_peak * 1e12) * 1e3 for k in ("fwd", "bwd")}       # ms
            t_fma = {k: 2.0 * fm_[k] * B * n_steps * 4 / (ffma_peak * 1e12) * 1e3 for k in ("fwd", "bwd")}
            bound_ms = max(t_tensor["bwd"], t_fma["bwd"])
            issued = 2.0 * tm_["bwd"] * evals / (bwd_ms * 1e-3) / 1e12
            roofline = {
                "bound": "tensor", "kernel": "rk4_bwd_tc_kernel<1,8> (tcgen05, fp32 through 3-way bf16 splits)",
                "achieved": issued, "peak": t_peak, "unit": "TFLOP/s", "frac": bound_ms / bwd_ms, "peak_source": t_peak_src,
                "note": "achieved = bf16 flops issued to the tensor pipe by the adjoint (6 bf16 products per fp32 product in the layer and "
                        "delta chains, 3 in the weight gradients) / its CUDA-event time; frac = mixed bound max(tensor-issued / bf16 peak, "
                        "FMA-pipe remainder / FFMA peak) / measured time",
                "mixed_bound_ms": {"tensor": t_tensor["bwd"], "fma": t_fma["bwd"], "measured": bwd_ms},
                "fwd_kernel": {"ms": fwd_ms, "frac": max(t_tensor["fwd"], t_fma["fwd"]) / fwd_ms,
                               "issued_tflops": 2.0 * tm_["fwd"] * evals / (fwd_ms * 1e-3) / 1e12},
            }
        else:
            roofline = {"bound": "fp32-fma", "kernel": "rk4_bwd_kernel (FFMA engine)", "achieved": ach_bwd, "peak": ffma_peak,
                        "unit": "TFLOP/s", "frac": ach_bwd / ffma_peak, "peak_source": "measured live: ude_measure_ffma_tflops",
                        "fwd_kernel": {"ms": fwd_ms, "frac": ach_fwd / ffma_peak}}
        roofline.update({
            "traffic": traffic, "algorithmic_bytes_per_launch": B * n_steps * (48 + 32) + B * 64, "bwd_ms": bwd_ms,
            # SURVEY 8d's definition (algorithmic fp32 flops against the fp32 FMA pipe), kept as the secondary figure
            "ffma_algorithmic": {"peak_tflops": ffma_peak, "peak_source": "measured live: ude_measure_ffma_tflops (register-resident FFMA chains)",
                                 "bwd_achieved_tflops": ach_bwd, "bwd_frac": ach_bwd / ffma_peak, "fwd_achieved_tflops": ach_fwd,
                                 "fwd_frac": ach_fwd / ffma_peak,
                                 "whole_step_frac": (B * n_steps * fl["total"] / ((fwd_ms + bwd_ms) * 1e-3) / 1e12) / ffma_peak},
            "hbm_gbs_algorithmic": B * n_steps * bytes_ts / (ms_per_step * 1e-3) / 1e9, "hbm_peak_gbs": peaks.get("hbm_gbs")})
        cpu = None
        secondary = None
        if world == 1:
            if not args.no_cpu_baseline:
                cpu = cpu_baseline_record(os.cpu_count() or 1)
            if not args.no_secondary:
                try:
                    secondary = secondary_configs(dev, ffma_peak, not args.no_cpu_baseline)
                except Exception as e:        # the headline line must not die with a secondary config
                    secondary = {"error": f"{type(e).__name__}: {e}"}
        tdq = "absent"
        try:
            import torchdiffeq
            tdq = torchdiffeq.__version__
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if not args.batch else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD + (", side sums + dtheta all-reduced over NCCL" if world > 1 else ""),
                       "trajectories_total": world * B, "trajectories_per_gpu": B, "grid_steps": n_steps,
                       "parallelism": f"trajectory-shard x{world}",
                       "l2": f"inputs larger than L2 (per pass and GPU: y0 {B * 32 >> 20} MiB + cotangent {B * 288 >> 20} MiB + tape {B * 384 >> 20} MiB)",
                       "tile_traj": int(plan.tile_traj), "ctas": int(plan.n_ctas), "torchdiffeq": tdq},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(y0_host.numel() * 4),
                    "d2h_bytes_per_step": int(res.numel() * 4), "ms_per_step": e2e_sec * 1e3,
                    "api": e2e_api,
                    "eager_odeint_decoded": {"value": world * B * n_steps / fused_sec, "ms_per_step": fused_sec * 1e3},
                    "cuda_graph": graphed,
                    "plain_odeint": {"value": world * B * n_steps / plain_sec, "ms_per_step": plain_sec * 1e3,
                                     "api": "ude_b200.odeint, then Decoder + losses as torch ops on the [T,B,R,L] trajectories (lib/VAE.py:137-138,186)"},
                    "fused_vs_plain_max_rel_diff": e2e_agree},
            "gpu_launches": 6 * args.steps,
            "clocks": clocks,
        }
        if weak is not None:
            line["weak_scaling"] = weak
        if check is not None:
            line["dist_check"] = check
        if secondary is not None:
            line["secondary"] = secondary
        print(json.dumps(line), flush=True)
    if world > 1:
        # The graphed end-to-end leg captured NCCL kernels; tearing the communicator down while such graphs exist can wait forever
        # (observed: the line above printed, then destroy_process_group() never returned).  Everything is measured and printed:
        # meet at a barrier, then leave without the teardown.
        dist.barrier()
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="trajectories PER GPU (default: 2**20 in total, sharded)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-graph-e2e", action="store_true", help="time the end-to-end leg eagerly only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
