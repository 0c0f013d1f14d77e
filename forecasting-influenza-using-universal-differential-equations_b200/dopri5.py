"""Adaptive Dopri5 front end — ``odeint(func, y0, t, rtol=, atol=)`` with torchdiffeq's default
method (the one the reference reaches at /root/reference/lib/train_functions.py:190).

Batch semantics.  torchdiffeq uses ONE step size for the whole batch (RMS error norm over the
entire tensor).  Two fused modes:
* shared step (``options['per_trajectory'] = False``, and the DEFAULT when the batch fits one
  resident tile per SM): torchdiffeq's semantics -- one norm, one dt, one accept / reject for
  all trajectories (cooperative launch, batch-wide norms summed across CTAs); the side channels
  cover EVERY evaluation (the initial-step probe and rejected attempts too), as the reference's
  python lists do.
* per trajectory (``options['per_trajectory'] = True``; the fallback, with a warning, for
  batches beyond the shared mode's capacity): one controller per trajectory, i.e. what
  torchdiffeq computes when called on each row alone (B == 1); side channels cover accepted
  stages only (1 + 6 * n_accept evaluations per trajectory).  See DESIGN.md.
The adjoint replays the accepted steps with their step sizes frozen (discrete adjoint of the
accepted-step map); side-channel cotangents act through accepted evaluations.

Extra (fused-path only) options: ``max_accepted_steps`` (tape capacity, default 256) and
``return_stats=True`` -> ``(y, stats)`` with per-trajectory accepted / rejected counts.
"""
from __future__ import annotations

import ctypes as C
import warnings

import torch

from . import _lib
from .funcs import FuncInfo
from .odeint import _ptr, _stream_ptr, install_side_channels

STATUS_TEXT = {1: "max_accepted_steps exceeded (raise options['max_accepted_steps'])",
               2: "max_num_steps exceeded", 3: "underflow in dt / non-finite state"}


class PerTrajectoryStepControl(UserWarning):
    """The fused Dopri5 was called on a batch without ``options['per_trajectory']`` being stated."""


def _opts(rtol, atol, options, batch: int = 1) -> "_lib.UdeDopri5Opts":
    known = {"first_step", "safety", "ifactor", "dfactor", "max_num_steps", "max_accepted_steps", "return_stats",
             "per_trajectory", "dtype", "process_group"}
    unknown = set(options) - known
    if unknown:
        raise ValueError(f"unexpected dopri5 options {sorted(unknown)}")
    o = _lib.UdeDopri5Opts()
    # per_trajectory: True -> one controller per row; False -> torchdiffeq's shared step (error if the batch does not fit);
    # absent -> shared step when it fits, else per trajectory with a warning (dopri5_solve)
    o.shared_step = 0 if (options.get("per_trajectory", None) is True or batch == 1) else 1
    o.rtol = float(rtol)
    o.atol = float(atol)
    fs = options.get("first_step", None)
    o.first_step = float(fs) if fs is not None else -1.0
    o.safety = float(options.get("safety", 0.9))
    o.ifactor = float(options.get("ifactor", 10.0))
    o.dfactor = float(options.get("dfactor", 0.2))
    o.max_steps = int(options.get("max_accepted_steps", 256))
    o.max_attempts = int(min(options.get("max_num_steps", 2 ** 31 - 1), 2 ** 31 - 1))
    return o


class _Dopri5Solve(torch.autograd.Function):
    last_steps = None

    @staticmethod
    def forward(ctx, theta, y0, info: FuncInfo, t_dev, opts, need_tape: bool):
        L = _lib.lib()
        dev = y0.device
        B, R, Ld = y0.shape
        n_out = int(t_dev.numel())
        desc = info.desc(B)
        plan = _lib.UdePlan()
        with torch.cuda.device(dev):
            _lib.check(L.ude_plan(C.byref(desc), _lib.METHOD["dopri5"], opts.max_steps, C.byref(plan)), "ude_plan")
            theta_c = theta.detach().contiguous()
            y0_c = y0.detach().contiguous()
            y_out = torch.empty((n_out, B, R, Ld), dtype=torch.float32, device=dev)
            side = torch.empty(8, dtype=torch.float64, device=dev)
            counts = torch.zeros((B, 4), dtype=torch.int32, device=dev)
            ws = torch.empty(int(plan.workspace_bytes), dtype=torch.uint8, device=dev)
            tape = torch.empty(int(plan.tape_bytes), dtype=torch.uint8, device=dev) if need_tape else None
            st = L.ude_dopri5_forward(C.byref(desc), C.byref(opts), _ptr(theta_c), _ptr(y0_c), n_out, _ptr(t_dev),
                                      _ptr(y_out), _ptr(tape), _ptr(counts), _ptr(side), _ptr(ws),
                                      int(plan.workspace_bytes), _stream_ptr(dev))
            _lib.check(st, "ude_dopri5_forward")
        ctx.info, ctx.opts, ctx.t_dev = info, opts, t_dev
        ctx.plan_ws = int(plan.workspace_bytes)
        ctx.has_tape = tape is not None
        ctx.save_for_backward(theta_c, y0_c, tape if tape is not None else torch.empty(0, device=dev), counts)
        ctx.mark_non_differentiable(counts)
        # accepted-step record (t0, dt) per trajectory, fp64: the tail of the tape (csrc/ude_host.cu: dopri_tape_x_bytes)
        ctx.steps = None
        if tape is not None:
            x_bytes = ((opts.max_steps + 1) * 6 * 3 * R * B * 4 + 255) // 256 * 256
            ctx.steps = tape[x_bytes:x_bytes + opts.max_steps * B * 16].view(torch.float64).view(opts.max_steps, B, 2)
        _Dopri5Solve.last_steps = ctx.steps
        return y_out, side, counts

    @staticmethod
    def backward(ctx, g_out, g_side, _g_counts):
        if not ctx.has_tape:
            raise RuntimeError("fused odeint was run without a tape (no input required grad)")
        L = _lib.lib()
        theta_c, y0_c, tape, counts = ctx.saved_tensors
        dev = y0_c.device
        desc = ctx.info.desc(y0_c.shape[0])
        n_out = int(ctx.t_dev.numel())
        with torch.cuda.device(dev):
            g_out = g_out.contiguous().float()
            gs = g_side[:5].to(torch.float32).contiguous() if g_side is not None else None
            g_y0 = torch.empty_like(y0_c)
            g_theta = torch.empty_like(theta_c)
            ws = torch.empty(ctx.plan_ws, dtype=torch.uint8, device=dev)
            st = L.ude_dopri5_backward(C.byref(desc), C.byref(ctx.opts), _ptr(theta_c), _ptr(y0_c), n_out,
                                       _ptr(ctx.t_dev), _ptr(tape), _ptr(counts), _ptr(g_out), _ptr(gs), _ptr(g_y0),
                                       _ptr(g_theta), _ptr(ws), ctx.plan_ws, _stream_ptr(dev))
            _lib.check(st, "ude_dopri5_backward")
        return g_theta, g_y0, None, None, None, None


class Dopri5Stats:
    """Per-trajectory controller statistics of the last fused solve."""

    def __init__(self, counts: torch.Tensor, steps=None):
        self.n_accept = counts[:, 0]
        self.n_reject = counts[:, 1]
        self.status = counts[:, 2]
        self.n_attempts = counts[:, 3]
        self.steps = steps          # [max_accepted_steps, B, 2] fp64 (t0, dt) of every accepted step (rows >= n_accept unused);
                                    # None when the solve kept no tape (no input required grad)


def dopri5_solve(func, info: FuncInfo, y0, t_cpu, rtol, atol, options):
    if info.bayes:
        raise NotImplementedError("Bayes_* funcs (weights re-drawn at every RHS evaluation) are fused for rk4, the reference's method")
    opts = _opts(rtol, atol, options, int(y0.shape[0]))
    t_dev = t_cpu.to(torch.float64).to(y0.device)
    theta = info.theta()
    need_tape = torch.is_grad_enabled() and (theta.requires_grad or y0.requires_grad)
    try:
        y_out, side, counts = _Dopri5Solve.apply(theta, y0, info, t_dev, opts, need_tape)
    except _lib.UdeError as e:
        if not (opts.shared_step and e.status == _lib.UDE_ERR_CAPACITY and "per_trajectory" not in options):
            raise
        # the default call must not silently differ from torchdiffeq's batch semantics: say so
        warnings.warn(
            f"ude_b200.odeint(method='dopri5'): a batch of {int(y0.shape[0])} trajectories exceeds the shared-step mode (one "
            "resident tile per SM); falling back to one controller PER TRAJECTORY (= torchdiffeq applied to every row alone). "
            "Results agree to the controller's tolerance, not step for step, and the side channels then cover accepted stages "
            "only.  Pass options={'per_trajectory': True} to acknowledge.", PerTrajectoryStepControl, stacklevel=3)
        opts.shared_step = 0
        y_out, side, counts = _Dopri5Solve.apply(theta, y0, info, t_dev, opts, need_tape)
    worst = int(counts[:, 2].max())          # torchdiffeq asserts on these conditions; so do we
    if worst != 0:
        raise _lib.UdeError("fused dopri5: " + STATUS_TEXT.get(worst, f"status {worst}"))
    group = options.get("process_group", None)
    if group is not None:
        from .dist import all_reduce_side
        side = all_reduce_side(side, None if group is True else group)
    install_side_channels(func, info, side)
    if options.get("return_stats", False):
        return y_out, Dopri5Stats(counts, _Dopri5Solve.last_steps)
    return y_out
