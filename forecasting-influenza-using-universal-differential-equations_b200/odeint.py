"""``odeint`` — drop-in for ``torchdiffeq.odeint`` on the reference's hot path.

    latent = odeint(self.ode, z, t, method='rk4', options=dict(step_size=h))
                                                   /root/reference/lib/VAE.py:137

Same signature and return value (``[len(t), *y0.shape]``); ``func`` must be an
``Fp`` / ``Fa`` / ``FaFp`` module (this package's or the reference's own
``lib.models`` classes) living on a CUDA device in fp32.  The whole
integration is ONE fused CUDA kernel launch; ``loss.backward()`` runs ONE fused
discrete-adjoint kernel and routes ``dL/dtheta`` into the module parameters and
``dL/dy0`` to whatever produced ``y0`` (the encoder).  Anything else raises —
there is no CPU path and no generic-callable path.

Side channels (lib/models.py:137,187,238,252): the reference appends beta/gamma
and Fa of *every* RHS evaluation to python lists that the loss later reduces
(``posterior()`` -> mean/std, lib/models.py:152-156; ``torch.norm(torch.stack(
tracker))``, lib/VAE.py:180).  The kernels emit the reductions instead; after a
solve ``func.params`` / ``func.tracker`` receive small differentiable tensors
for which those two expressions return exactly the reference's values.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from collections import OrderedDict
from typing import Optional

import numpy as np
import torch

from . import _lib
from .funcs import FuncInfo, UnsupportedFunc, describe_func

_GRID_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_GRID_CACHE_MAX = 64


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


# ---------------------------------------------------------------------------------------
# fixed grid tables (host side, fp32, bit-exact restatement of torchdiffeq — see ude_rk4_grid)
# ---------------------------------------------------------------------------------------
def rk4_grid_tables(t_host: np.ndarray, step_size: Optional[float]):
    """Returns (dt[n_steps] f32, out_step[T] i32, out_kind[T] i32, out_w[T] f32) as numpy arrays."""
    L = _lib.lib()
    t_host = np.ascontiguousarray(t_host, dtype=np.float32)
    T = t_host.shape[0]
    use = 0 if step_size is None else 1
    h = 0.0 if step_size is None else float(np.float32(step_size))
    cap = 1024
    while True:
        dt = np.empty(cap, dtype=np.float32)
        n_steps = C.c_int32(0)
        out_step = np.empty(T, dtype=np.int32)
        out_kind = np.empty(T, dtype=np.int32)
        out_w = np.empty(T, dtype=np.float32)
        st = L.ude_rk4_grid(t_host.ctypes.data, T, h, use, cap, dt.ctypes.data, C.byref(n_steps),
                            out_step.ctypes.data, out_kind.ctypes.data, out_w.ctypes.data)
        if st == _lib.UDE_ERR_CAPACITY:
            cap = int(n_steps.value)
            continue
        _lib.check(st, "ude_rk4_grid")
        return dt[: n_steps.value].copy(), out_step, out_kind, out_w


def _grid_on_device(t: torch.Tensor, step_size, device):
    t_host = t.detach().to("cpu", torch.float32).numpy()
    if step_size is None:
        h = None
    elif isinstance(step_size, torch.Tensor):
        h = float(step_size.detach().to("cpu", torch.float32))
    else:
        h = float(np.float32(step_size))
    key = (t_host.tobytes(), h, str(device))
    hit = _GRID_CACHE.get(key)
    if hit is not None:
        _GRID_CACHE.move_to_end(key)
        return hit
    dt, out_step, out_kind, out_w = rk4_grid_tables(t_host, h)
    dev = tuple(torch.from_numpy(a).to(device) for a in (dt, out_step, out_kind, out_w))
    val = (dev, int(dt.shape[0]))
    _GRID_CACHE[key] = val
    if len(_GRID_CACHE) > _GRID_CACHE_MAX:
        _GRID_CACHE.popitem(last=False)
    return val


# ---------------------------------------------------------------------------------------
# autograd bridge
# ---------------------------------------------------------------------------------------
def flat_theta(params) -> torch.Tensor:
    """The C ABI's flat weight vector from the module's (W, b) tensors in theta order.  Zero-copy when they already are
    consecutive views of ONE buffer (``ude_b200.flatten_parameters``), else one ``torch.cat``."""
    p0 = params[0]
    ptr, esz = p0.data_ptr(), p0.element_size()
    for p in params:
        if p.data_ptr() != ptr or not p.is_contiguous():
            return torch.cat([q.detach().reshape(-1) for q in params])
        ptr += p.numel() * esz
    n = (ptr - p0.data_ptr()) // esz
    base = p0.detach()
    return base.as_strided((n,), (1,))      # stays inside the shared buffer by construction


def split_like(flat: torch.Tensor, params):
    """Views of `flat` shaped like `params` (gradients come back as views of ONE buffer: no per-parameter kernels)."""
    out, off = [], 0
    for p in params:
        n = p.numel()
        out.append(flat[off:off + n].view(p.shape))
        off += n
    return out


def flatten_parameters(func) -> torch.Tensor:
    """Re-home the ODE func's Linear weights and biases as consecutive views of one flat buffer (theta order), so that a
    solve hands the buffer to the kernels without gathering it and the optimizer can treat it as one tensor.
    Returns the buffer; the module's parameters keep their names, shapes and values (state_dict unchanged)."""
    info = describe_func(func)
    ps = [p for l in info.linears() for p in (l.weight, l.bias)]
    flat = torch.cat([p.detach().reshape(-1) for p in ps])
    for p, v in zip(ps, split_like(flat, ps)):
        p.data = v
    return flat


class _Rk4Solve(torch.autograd.Function):
    """(y0, *params) -> (y_out [T,B,R,L], side [8] fp64).  One fused launch each way."""

    @staticmethod
    def forward(ctx, y0, info: FuncInfo, tables, n_steps: int, n_out: int, need_tape: bool, *params):
        L = _lib.lib()
        dev = y0.device
        B, R, Ld = y0.shape
        desc = info.desc(B)
        plan = _lib.UdePlan()
        with torch.cuda.device(dev):
            _lib.check(L.ude_plan(C.byref(desc), _lib.METHOD["rk4"], n_steps, C.byref(plan)), "ude_plan")
            theta_c = flat_theta(params)
            y0_c = y0.detach().contiguous()
            y_out = torch.empty((n_out, B, R, Ld), dtype=torch.float32, device=dev)
            side = torch.empty(8, dtype=torch.float64, device=dev)
            ws = torch.empty(int(plan.workspace_bytes), dtype=torch.uint8, device=dev)
            tape = torch.empty(int(plan.tape_bytes), dtype=torch.uint8, device=dev) if need_tape else None
            dt, out_step, out_kind, out_w = tables
            st = L.ude_rk4_forward(C.byref(desc), _ptr(theta_c), _ptr(y0_c), n_steps, _ptr(dt), n_out,
                                   _ptr(out_step), _ptr(out_kind), _ptr(out_w), _ptr(y_out), _ptr(tape), _ptr(side),
                                   _ptr(ws), int(plan.workspace_bytes), _stream_ptr(dev))
            _lib.check(st, "ude_rk4_forward")
        ctx.info, ctx.tables, ctx.n_steps, ctx.n_out = info, tables, n_steps, n_out
        ctx.plan_ws = int(plan.workspace_bytes)
        # the workspace travels to the backward: it still holds THIS theta's packed weight images, which the adjoint then
        # does not build again (UDE_FLAG_WEIGHTS_PACKED); theta itself is a snapshot when it was gathered with cat
        ctx.save_for_backward(theta_c, y0_c, tape if tape is not None else torch.empty(0, device=dev),
                              ws if tape is not None else torch.empty(0, device=dev))
        ctx.has_tape = tape is not None
        ctx.param_meta = [(p.shape, p.numel()) for p in params]
        return y_out, side

    @staticmethod
    def backward(ctx, g_out, g_side):
        if not ctx.has_tape:
            raise RuntimeError("fused odeint was run without a tape (no input required grad)")
        L = _lib.lib()
        theta_c, y0_c, tape, ws = ctx.saved_tensors
        info: FuncInfo = ctx.info
        dev = y0_c.device
        B = y0_c.shape[0]
        desc = info.desc(B)
        desc.flags |= _lib.FLAG_WEIGHTS_PACKED
        with torch.cuda.device(dev):
            if g_out is None:
                g_out = torch.zeros((ctx.n_out, *y0_c.shape), dtype=torch.float32, device=dev)
            g_out = g_out.contiguous().float()
            gs = None
            if g_side is not None:
                gs = g_side[:5].to(torch.float32).contiguous()
            g_y0 = torch.empty_like(y0_c)
            g_theta = torch.empty(theta_c.numel(), dtype=torch.float32, device=dev)
            dt, out_step, out_kind, out_w = ctx.tables
            st = L.ude_rk4_backward(C.byref(desc), _ptr(theta_c), _ptr(y0_c), ctx.n_steps, _ptr(dt), ctx.n_out,
                                    _ptr(out_step), _ptr(out_kind), _ptr(out_w), _ptr(tape), _ptr(g_out), _ptr(gs),
                                    _ptr(g_y0), _ptr(g_theta), _ptr(ws), ctx.plan_ws, _stream_ptr(dev))
            _lib.check(st, "ude_rk4_backward")
        grads, off = [], 0
        for shape, n in ctx.param_meta:
            grads.append(g_theta[off:off + n].view(shape))
            off += n
        return (g_y0, None, None, None, None, None, *grads)


class _Rk4SolveReparam(torch.autograd.Function):
    """(eps, std, mean, *params) -> (y_out [T, S*b, R, L], side): the y0 rows are produced in the solve's prologue and the
    adjoint's epilogue hands back dL/dz (include/ude_b200.h UdeReparam)."""

    @staticmethod
    def forward(ctx, eps, std, mean, info: FuncInfo, tables, n_steps: int, n_out: int, need_tape: bool, *params):
        L = _lib.lib()
        dev = eps.device
        S, b, R, Lm1 = eps.shape
        B, Ld = S * b, Lm1 + 1
        desc = info.desc(B)
        plan = _lib.UdePlan()
        with torch.cuda.device(dev):
            _lib.check(L.ude_plan(C.byref(desc), _lib.METHOD["rk4"], n_steps, C.byref(plan)), "ude_plan")
            theta_c = flat_theta(params)
            eps_c, std_c, mean_c = eps.detach().contiguous(), std.detach().contiguous(), mean.detach().contiguous()
            rp = _lib.UdeReparam(eps_c.data_ptr(), mean_c.data_ptr(), std_c.data_ptr(), S, b)
            y_out = torch.empty((n_out, B, R, Ld), dtype=torch.float32, device=dev)
            side = torch.empty(8, dtype=torch.float64, device=dev)
            ws = torch.empty(int(plan.workspace_bytes), dtype=torch.uint8, device=dev)
            tape = torch.empty(int(plan.tape_bytes), dtype=torch.uint8, device=dev) if need_tape else None
            dt, out_step, out_kind, out_w = tables
            st = L.ude_rk4_forward_reparam(C.byref(desc), _ptr(theta_c), C.byref(rp), n_steps, _ptr(dt), n_out, _ptr(out_step),
                                           _ptr(out_kind), _ptr(out_w), _ptr(y_out), _ptr(tape), _ptr(side), _ptr(ws),
                                           int(plan.workspace_bytes), _stream_ptr(dev))
            _lib.check(st, "ude_rk4_forward_reparam")
        ctx.info, ctx.tables, ctx.n_steps, ctx.n_out = info, tables, n_steps, n_out
        ctx.plan_ws = int(plan.workspace_bytes)
        empty = torch.empty(0, device=dev)
        # y0 for the adjoint is output 0 itself (a view: nothing is copied)
        ctx.save_for_backward(theta_c, eps_c, std_c, mean_c, y_out[0] if need_tape else empty, tape if need_tape else empty,
                              ws if need_tape else empty)
        ctx.has_tape = need_tape
        ctx.param_meta = [(p.shape, p.numel()) for p in params]
        return y_out, side

    @staticmethod
    def backward(ctx, g_out, g_side):
        if not ctx.has_tape:
            raise RuntimeError("fused odeint was run without a tape (no input required grad)")
        L = _lib.lib()
        theta_c, eps_c, std_c, mean_c, y0_c, tape, ws = ctx.saved_tensors
        S, b, R, Lm1 = eps_c.shape
        dev = eps_c.device
        desc = ctx.info.desc(S * b)
        desc.flags |= _lib.FLAG_WEIGHTS_PACKED
        with torch.cuda.device(dev):
            if g_out is None:
                g_out = torch.zeros((ctx.n_out, S * b, R, Lm1 + 1), dtype=torch.float32, device=dev)
            g_out = g_out.contiguous().float()
            gs = g_side[:5].to(torch.float32).contiguous() if g_side is not None else None
            g_y0 = torch.empty((S * b, R, Lm1 + 1), dtype=torch.float32, device=dev)
            g_z = torch.empty((S, b, R, Lm1), dtype=torch.float32, device=dev)
            g_theta = torch.empty(theta_c.numel(), dtype=torch.float32, device=dev)
            rp = _lib.UdeReparam(eps_c.data_ptr(), mean_c.data_ptr(), std_c.data_ptr(), S, b)
            dt, out_step, out_kind, out_w = ctx.tables
            st = L.ude_rk4_backward_reparam(C.byref(desc), _ptr(theta_c), C.byref(rp), _ptr(y0_c), ctx.n_steps, _ptr(dt), ctx.n_out,
                                            _ptr(out_step), _ptr(out_kind), _ptr(out_w), _ptr(tape), _ptr(g_out), _ptr(gs),
                                            _ptr(g_y0), _ptr(g_z), _ptr(g_theta), _ptr(ws), ctx.plan_ws, _stream_ptr(dev))
            _lib.check(st, "ude_rk4_backward_reparam")
        # the sample-axis sums (tiny next to the solve): z = eps * std + mean
        g_eps = g_z * std_c if ctx.needs_input_grad[0] else None
        g_std = (g_z * eps_c).sum(0) if ctx.needs_input_grad[1] else None
        g_mean = g_z.sum(0) if ctx.needs_input_grad[2] else None
        grads, off = [], 0
        for shape, n in ctx.param_meta:
            grads.append(g_theta[off:off + n].view(shape))
            off += n
        return (g_eps, g_std, g_mean, None, None, None, None, None, *grads)


class _Rk4SolveBayes(torch.autograd.Function):
    """(y0, z [4 * n_steps, n_theta]; means..., stds...) -> (y_out, side): Bayesian ODE funcs.  Evaluation e of the
    solve runs with theta_e = mean + z_e * |std| (models_bayes.py:44-48); the adjoint returns dL/dmean = sum_e dL/dtheta_e and
    dL/dstd = sign(std) * sum_e dL/dtheta_e * z_e."""

    @staticmethod
    def forward(ctx, y0, z, info: FuncInfo, tables, n_steps: int, n_out: int, need_tape: bool, n_par: int, *params):
        L = _lib.lib()
        dev = y0.device
        B, R, Ld = y0.shape
        desc = info.desc(B)
        desc.flags |= _lib.FLAG_BAYES
        plan = _lib.UdePlan()
        with torch.cuda.device(dev):
            _lib.check(L.ude_plan(C.byref(desc), _lib.METHOD["rk4"], n_steps, C.byref(plan)), "ude_plan (bayes)")
            mean = torch.cat([p.detach().reshape(-1) for p in params[:n_par]])
            std = torch.cat([p.detach().reshape(-1) for p in params[n_par:]])
            z_c = z.detach().contiguous()
            theta_e = torch.addcmul(mean, z_c, std.abs())              # [4 * n_steps, n_theta]
            y0_c = y0.detach().contiguous()
            y_out = torch.empty((n_out, B, R, Ld), dtype=torch.float32, device=dev)
            side = torch.empty(8, dtype=torch.float64, device=dev)
            ws = torch.empty(int(plan.workspace_bytes), dtype=torch.uint8, device=dev)
            tape = torch.empty(int(plan.tape_bytes), dtype=torch.uint8, device=dev) if need_tape else None
            dt, out_step, out_kind, out_w = tables
            st = L.ude_rk4_forward_bayes(C.byref(desc), _ptr(theta_e), _ptr(y0_c), n_steps, _ptr(dt), n_out, _ptr(out_step),
                                         _ptr(out_kind), _ptr(out_w), _ptr(y_out), _ptr(tape), _ptr(side), _ptr(ws),
                                         int(plan.workspace_bytes), _stream_ptr(dev))
            _lib.check(st, "ude_rk4_forward_bayes")
        ctx.info, ctx.tables, ctx.n_steps, ctx.n_out, ctx.n_par = info, tables, n_steps, n_out, n_par
        ctx.plan_ws = int(plan.workspace_bytes)
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(theta_e if need_tape else empty, z_c, std, y0_c, tape if need_tape else empty, ws if need_tape else empty)
        ctx.has_tape = need_tape
        ctx.param_meta = [(p.shape, p.numel()) for p in params[:n_par]]
        return y_out, side

    @staticmethod
    def backward(ctx, g_out, g_side):
        if not ctx.has_tape:
            raise RuntimeError("fused odeint was run without a tape (no input required grad)")
        L = _lib.lib()
        theta_e, z_c, std, y0_c, tape, ws = ctx.saved_tensors
        dev = y0_c.device
        desc = ctx.info.desc(y0_c.shape[0])
        desc.flags |= _lib.FLAG_BAYES | _lib.FLAG_WEIGHTS_PACKED
        with torch.cuda.device(dev):
            if g_out is None:
                g_out = torch.zeros((ctx.n_out, *y0_c.shape), dtype=torch.float32, device=dev)
            g_out = g_out.contiguous().float()
            gs = g_side[:5].to(torch.float32).contiguous() if g_side is not None else None
            g_y0 = torch.empty_like(y0_c)
            g2 = torch.empty((2, std.numel()), dtype=torch.float32, device=dev)
            dt, out_step, out_kind, out_w = ctx.tables
            st = L.ude_rk4_backward_bayes(C.byref(desc), _ptr(theta_e), _ptr(z_c), _ptr(y0_c), ctx.n_steps, _ptr(dt), ctx.n_out,
                                          _ptr(out_step), _ptr(out_kind), _ptr(out_w), _ptr(tape), _ptr(g_out), _ptr(gs), _ptr(g_y0),
                                          _ptr(g2), _ptr(ws), ctx.plan_ws, _stream_ptr(dev))
            _lib.check(st, "ude_rk4_backward_bayes")
        g_std = g2[1] * torch.sign(std)
        gm, gsd, off = [], [], 0
        for shape, n in ctx.param_meta:
            gm.append(g2[0, off:off + n].view(shape))
            gsd.append(g_std[off:off + n].view(shape))
            off += n
        return (g_y0, None, None, None, None, None, None, None, *gm, *gsd)


class _Rk4SolveDecoded(torch.autograd.Function):
    """(y0 | eps, std, mean; dec_w, dec_b; *params) -> (y_pred [T, B, n_dec], side [8] fp64 with side[7] = the
    latent_init_loss sum): the decoder and the reg loss are applied where the outputs are produced, the trajectories are
    never written (include/ude_b200.h UdeDecoder)."""

    @staticmethod
    def forward(ctx, y0, eps, std, mean, dec_w, dec_b, info: FuncInfo, tables, n_steps: int, n_out: int, need_tape: bool,
                bayes_z, n_par: int, *params):
        # Bayes funcs (bayes_z given): params = (*means, *stds), n_par of each; evaluation e runs with mean + z_e * |std|
        L = _lib.lib()
        reparam = y0 is None
        bayes = bayes_z is not None
        dev = dec_w.device
        if reparam:
            S, b, R, Lm1 = eps.shape
            B, Ld = S * b, Lm1 + 1
        else:
            B, R, Ld = y0.shape
        n_dec = dec_w.shape[0]
        desc = info.desc(B)
        if bayes:
            desc.flags |= _lib.FLAG_BAYES
        if n_dec != 1 and not (desc.flags & 8):
            desc.flags |= 4                          # wider decoders run the FFMA engine: size the plan for it
        plan = _lib.UdePlan()
        with torch.cuda.device(dev):
            _lib.check(L.ude_plan(C.byref(desc), _lib.METHOD["rk4"], n_steps, C.byref(plan)), "ude_plan")
            if bayes:
                std_flat = torch.cat([p.detach().reshape(-1) for p in params[n_par:]])
                z_c = bayes_z.detach().contiguous()
                theta_c = torch.addcmul(torch.cat([p.detach().reshape(-1) for p in params[:n_par]]), z_c, std_flat.abs())
            else:
                std_flat = z_c = torch.empty(0, device=dev)
                theta_c = flat_theta(params)
            w_c, b_c = dec_w.detach().contiguous(), dec_b.detach().contiguous()
            dec = _lib.UdeDecoder(w_c.data_ptr(), b_c.data_ptr(), n_dec)
            rp = None
            if reparam:
                eps_c, std_c, mean_c = eps.detach().contiguous(), std.detach().contiguous(), mean.detach().contiguous()
                rp = _lib.UdeReparam(eps_c.data_ptr(), mean_c.data_ptr(), std_c.data_ptr(), S, b)
                y0_c = torch.empty((B, R, Ld), dtype=torch.float32, device=dev)      # produced by the prologue
            else:
                y0_c = y0.detach().contiguous()
                eps_c = std_c = mean_c = torch.empty(0, device=dev)
            y_pred = torch.empty((n_out, B, n_dec), dtype=torch.float32, device=dev)
            mask = torch.empty((n_out, B, R), dtype=torch.uint8, device=dev)
            side = torch.empty(8, dtype=torch.float64, device=dev)
            ws = torch.empty(int(plan.workspace_bytes), dtype=torch.uint8, device=dev)
            tape = torch.empty(int(plan.tape_bytes), dtype=torch.uint8, device=dev) if need_tape else None
            dt, out_step, out_kind, out_w = tables
            fwd = L.ude_rk4_forward_bayes_decoded if bayes else L.ude_rk4_forward_decoded
            st = fwd(C.byref(desc), _ptr(theta_c), _ptr(None if reparam else y0_c),
                     C.byref(rp) if reparam else None, n_steps, _ptr(dt), n_out, _ptr(out_step),
                     _ptr(out_kind), _ptr(out_w), C.byref(dec), _ptr(y_pred), _ptr(mask),
                     _ptr(y0_c if reparam else None), _ptr(tape), _ptr(side), _ptr(ws),
                     int(plan.workspace_bytes), _stream_ptr(dev))
            _lib.check(st, "ude_rk4_forward_bayes_decoded" if bayes else "ude_rk4_forward_decoded")
        ctx.info, ctx.tables, ctx.n_steps, ctx.n_out, ctx.reparam, ctx.bayes = info, tables, n_steps, n_out, reparam, bayes
        ctx.plan_ws, ctx.n_ctas, ctx.desc_flags = int(plan.workspace_bytes), int(plan.n_ctas), int(desc.flags)
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(theta_c if (need_tape or not bayes) else empty, y0_c, eps_c, std_c, mean_c, w_c, b_c,
                              mask if need_tape else empty, tape if need_tape else empty, ws if need_tape else empty, z_c, std_flat)
        ctx.has_tape = need_tape
        ctx.param_meta = [(p.shape, p.numel()) for p in (params[:n_par] if bayes else params)]
        return y_pred, side

    @staticmethod
    def backward(ctx, g_pred, g_side):
        if not ctx.has_tape:
            raise RuntimeError("fused odeint was run without a tape (no input required grad)")
        L = _lib.lib()
        theta_c, y0_c, eps_c, std_c, mean_c, w_c, b_c, mask, tape, ws, z_c, std_flat = ctx.saved_tensors
        dev = w_c.device
        B, R, Ld = y0_c.shape
        n_dec = w_c.shape[0]
        desc = ctx.info.desc(B)
        desc.flags = ctx.desc_flags | _lib.FLAG_WEIGHTS_PACKED
        with torch.cuda.device(dev):
            if g_pred is None:
                g_pred = torch.zeros((ctx.n_out, B, n_dec), dtype=torch.float32, device=dev)
            g_pred = g_pred.contiguous().float()
            gs = None
            if g_side is not None:      # [0..4]: the usual side sums, [5]: dL/d(reg sum) = dL/dside[7]
                gs = torch.cat([g_side[:5], g_side[7:8]]).to(torch.float32).contiguous()
            g_y0 = torch.empty_like(y0_c)
            g_theta = torch.empty((2, std_flat.numel()) if ctx.bayes else theta_c.numel(), dtype=torch.float32, device=dev)
            n_d = n_dec * (3 * R + 1)
            g_dec = torch.empty(n_d, dtype=torch.float32, device=dev)
            partial = torch.empty(ctx.n_ctas * n_d, dtype=torch.float32, device=dev)
            dec = _lib.UdeDecoder(w_c.data_ptr(), b_c.data_ptr(), n_dec)
            rp, g_z = None, None
            if ctx.reparam:
                S, b = eps_c.shape[0], eps_c.shape[1]
                rp = _lib.UdeReparam(eps_c.data_ptr(), mean_c.data_ptr(), std_c.data_ptr(), S, b)
                g_z = torch.empty((S, b, R, Ld - 1), dtype=torch.float32, device=dev)
            dt, out_step, out_kind, out_w = ctx.tables
            if ctx.bayes:
                st = L.ude_rk4_backward_bayes_decoded(C.byref(desc), _ptr(theta_c), _ptr(z_c), _ptr(y0_c),
                                                      C.byref(rp) if rp is not None else None, ctx.n_steps, _ptr(dt), ctx.n_out,
                                                      _ptr(out_step), _ptr(out_kind), _ptr(out_w), C.byref(dec), _ptr(tape), _ptr(mask),
                                                      _ptr(g_pred), _ptr(gs), _ptr(g_y0), _ptr(g_z), _ptr(g_theta), _ptr(g_dec),
                                                      _ptr(partial), _ptr(ws), ctx.plan_ws, _stream_ptr(dev))
            else:
                st = L.ude_rk4_backward_decoded(C.byref(desc), _ptr(theta_c), _ptr(y0_c), C.byref(rp) if rp is not None else None,
                                                ctx.n_steps, _ptr(dt), ctx.n_out, _ptr(out_step), _ptr(out_kind), _ptr(out_w),
                                                C.byref(dec), _ptr(tape), _ptr(mask), _ptr(g_pred), _ptr(gs), _ptr(g_y0), _ptr(g_z),
                                                _ptr(g_theta), _ptr(g_dec), _ptr(partial), _ptr(ws), ctx.plan_ws, _stream_ptr(dev))
            _lib.check(st, "ude_rk4_backward_bayes_decoded" if ctx.bayes else "ude_rk4_backward_decoded")
        g_in = [None, None, None, None]
        if ctx.reparam:
            g_in[1] = g_z * std_c if ctx.needs_input_grad[1] else None
            g_in[2] = (g_z * eps_c).sum(0) if ctx.needs_input_grad[2] else None
            g_in[3] = g_z.sum(0) if ctx.needs_input_grad[3] else None
        else:
            g_in[0] = g_y0
        grads, off = [], 0
        if ctx.bayes:       # dL/dmean = sum_e dL/dtheta_e, dL/dstd = sign(std) * sum_e dL/dtheta_e * z_e
            g_std = g_theta[1] * torch.sign(std_flat)
            g_stds = []
            for shape, n in ctx.param_meta:
                grads.append(g_theta[0, off:off + n].view(shape))
                g_stds.append(g_std[off:off + n].view(shape))
                off += n
            grads += g_stds
        else:
            for shape, n in ctx.param_meta:
                grads.append(g_theta[off:off + n].view(shape))
                off += n
        return (*g_in, g_dec[:n_dec * 3 * R].view(n_dec, 3 * R), g_dec[n_dec * 3 * R:], None, None, None, None, None, None, None, *grads)


class _Rk4SolveGroups(torch.autograd.Function):
    """G independent weight sets in one call (C ABI ``n_groups = G``, group-major layouts):
    (theta [G, n_theta], y0 [G, B_g, R, L]) -> (y_out [G, T, B_g, R, L], side [G, 8] fp64)."""

    @staticmethod
    def forward(ctx, theta, y0, info: FuncInfo, tables, n_steps: int, n_out: int, need_tape: bool):
        L = _lib.lib()
        dev = y0.device
        G, Bg, R, Ld = y0.shape
        desc = info.desc(G * Bg, n_groups=G)
        plan = _lib.UdePlan()
        with torch.cuda.device(dev):
            _lib.check(L.ude_plan(C.byref(desc), _lib.METHOD["rk4"], n_steps, C.byref(plan)), "ude_plan")
            theta_c = theta.detach().contiguous()
            y0_c = y0.detach().contiguous()
            y_out = torch.empty((G, n_out, Bg, R, Ld), dtype=torch.float32, device=dev)
            side = torch.empty((G, 8), dtype=torch.float64, device=dev)
            ws = torch.empty(int(plan.workspace_bytes), dtype=torch.uint8, device=dev)
            tape = torch.empty(int(plan.tape_bytes), dtype=torch.uint8, device=dev) if need_tape else None
            dt, out_step, out_kind, out_w = tables
            st = L.ude_rk4_forward(C.byref(desc), _ptr(theta_c), _ptr(y0_c), n_steps, _ptr(dt), n_out,
                                   _ptr(out_step), _ptr(out_kind), _ptr(out_w), _ptr(y_out), _ptr(tape), _ptr(side),
                                   _ptr(ws), int(plan.workspace_bytes), _stream_ptr(dev))
            _lib.check(st, "ude_rk4_forward (groups)")
        ctx.info, ctx.tables, ctx.n_steps, ctx.n_out = info, tables, n_steps, n_out
        ctx.plan_ws = int(plan.workspace_bytes)
        ctx.save_for_backward(theta_c, y0_c, tape if tape is not None else torch.empty(0, device=dev))
        ctx.has_tape = tape is not None
        return y_out, side

    @staticmethod
    def backward(ctx, g_out, g_side):
        if not ctx.has_tape:
            raise RuntimeError("fused odeint was run without a tape (no input required grad)")
        L = _lib.lib()
        theta_c, y0_c, tape = ctx.saved_tensors
        info: FuncInfo = ctx.info
        dev = y0_c.device
        G, Bg = y0_c.shape[0], y0_c.shape[1]
        desc = info.desc(G * Bg, n_groups=G)
        with torch.cuda.device(dev):
            if g_out is None:
                g_out = torch.zeros((G, ctx.n_out, *y0_c.shape[1:]), dtype=torch.float32, device=dev)
            g_out = g_out.contiguous().float()
            gs = None
            if g_side is not None:
                gs = g_side[:, :5].to(torch.float32).contiguous()
            g_y0 = torch.empty_like(y0_c)
            g_theta = torch.empty_like(theta_c)
            ws = torch.empty(ctx.plan_ws, dtype=torch.uint8, device=dev)
            dt, out_step, out_kind, out_w = ctx.tables
            st = L.ude_rk4_backward(C.byref(desc), _ptr(theta_c), _ptr(y0_c), ctx.n_steps, _ptr(dt), ctx.n_out,
                                    _ptr(out_step), _ptr(out_kind), _ptr(out_w), _ptr(tape), _ptr(g_out), _ptr(gs),
                                    _ptr(g_y0), _ptr(g_theta), _ptr(ws), ctx.plan_ws, _stream_ptr(dev))
            _lib.check(st, "ude_rk4_backward (groups)")
        return g_theta, g_y0, None, None, None, None, None


# ---------------------------------------------------------------------------------------
# side channels -> the reference's python-list protocol
# ---------------------------------------------------------------------------------------
def _safe_sqrt(v: torch.Tensor) -> torch.Tensor:
    """sqrt with a finite (zero) gradient at 0, like torch.norm / std of a constant: sqrt'(0) = inf would turn a zero
    cotangent into NaN."""
    pos = v > 0
    return torch.where(pos, torch.sqrt(torch.where(pos, v, torch.ones_like(v))), torch.zeros_like(v))


def posterior_moments(side: torch.Tensor):
    """(mean [2], unbiased std [2]) of every (beta, gamma) row behind the side sums (one solve: side [8]; several: the SUM of
    their side vectors -- sums, sums of squares and N pool exactly)."""
    n = side[5]
    mean = side[0:2] / n
    var = (side[2:4] - n * mean * mean) / (n - 1.0)
    return mean, _safe_sqrt(torch.clamp(var, min=0.0))


def posterior_rows(side: torch.Tensor) -> torch.Tensor:
    """A [2, 2] tensor whose column mean / unbiased std equal the mean / std of every
    (beta, gamma) row the reference would have stacked (lib/models.py:152-156)."""
    mean, std = posterior_moments(side)
    half = std / math.sqrt(2.0)
    return torch.stack([mean - half, mean + half]).to(torch.float32)


def install_side_channels(func, info: FuncInfo, side: torch.Tensor) -> None:
    """Hand the in-kernel reductions to the reference's list protocol.  Reference-owned modules (lib.models.*) get ONE surrogate
    entry per solve, exact for `posterior()` / `torch.norm(torch.stack(tracker))` with one solve between `clear_tracking()`
    and the read (what lib/VAE.py:121,173,180 does); this package's own classes additionally keep the raw sums
    (`_side_sums`) and pool any number of solves exactly in their `posterior()`."""
    if info.rate_layers and isinstance(getattr(func, "params", None), list):
        func.params.append(posterior_rows(side))
        if isinstance(getattr(func, "_side_sums", None), list):
            func._side_sums.append(side)
    if info.aug_layers and isinstance(getattr(func, "tracker", None), list):
        func.tracker.append(_safe_sqrt(side[4]).to(torch.float32))         # ||stack of per-solve norms|| pools exactly


# ---------------------------------------------------------------------------------------
# public entry point
# ---------------------------------------------------------------------------------------
def _check_common(func, y0, t, method, event_fn):
    if event_fn is not None:
        raise NotImplementedError("event handling is not on the fused path")
    if not isinstance(y0, torch.Tensor):
        raise NotImplementedError("tuple-valued states are not on the fused path")
    if not torch.is_floating_point(y0):
        raise TypeError("`y0` must be a floating point Tensor")
    if not isinstance(t, torch.Tensor) or t.dim() != 1 or not torch.is_floating_point(t):
        raise ValueError("t must be a one dimensional floating point tensor")
    if t.numel() < 2:
        raise ValueError("t must hold at least two time points")
    info = describe_func(func)
    if not y0.is_cuda:
        raise RuntimeError("the fused UDE solver runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
    if y0.dtype != torch.float32:
        raise TypeError("the fused UDE solver is fp32 (as the reference, run_ode.py:30)")
    if y0.dim() != 3 or y0.shape[1] != info.n_regions or y0.shape[2] != info.latent_dim:
        raise ValueError(f"y0 must be [B, {info.n_regions}, {info.latent_dim}], got {tuple(y0.shape)}")
    for l in info.linears():
        if l.weight.device != y0.device or l.weight.dtype != torch.float32:
            raise RuntimeError("ODE func parameters must be fp32 on y0's device")
    return info


def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None):
    """torchdiffeq.odeint signature; fused sm_100a implementation for UDE funcs."""
    if method is None:
        method = "dopri5"
    if method not in ("rk4", "dopri5"):
        raise ValueError(f'Invalid method "{method}". The fused path implements rk4 and dopri5.')
    info = _check_common(func, y0, t, method, event_fn)
    options = {} if options is None else dict(options)
    t_cpu = t.detach().to("cpu")
    inc = bool((t_cpu[1:] > t_cpu[:-1]).all())
    if not inc:
        if bool((t_cpu[1:] < t_cpu[:-1]).all()):
            raise NotImplementedError("decreasing t is not on the reference's path")
        raise ValueError("t must be strictly increasing or decreasing")
    if method == "dopri5":
        from .dopri5 import dopri5_solve
        return dopri5_solve(func, info, y0, t_cpu, rtol, atol, options)

    group = options.pop("process_group", None)        # fused-path extra: all-reduce the side sums over this group
    info.kernels = options.pop("kernels", os.environ.get("UDE_B200_KERNELS") or None)   # fused-path extra: pin the kernel family
    bayes_z = options.pop("bayes_z", None)            # fused-path extra: the weight noise of a Bayes_* func, [4 * n_steps, n_theta]
    if bayes_z is not None and not info.bayes:
        raise ValueError("options['bayes_z'] goes with a Bayes_* func")
    unknown = set(options) - {"step_size", "interp", "perturb", "grid_constructor"}
    if unknown:
        raise ValueError(f"unexpected rk4 options {sorted(unknown)}")
    if options.get("interp", "linear") != "linear" or options.get("perturb", False) or \
            options.get("grid_constructor", None) is not None:
        raise NotImplementedError("only step_size grids with linear interpolation are fused")
    (tables, n_steps) = _grid_on_device(t_cpu, options.get("step_size", None), y0.device)
    params = [p for l in info.linears() for p in (l.weight, l.bias)]
    if info.bayes:
        # Bayes_* funcs draw new weights at every RHS evaluation (models_bayes.py:30-48).  The noise of all 4 * n_steps
        # evaluations comes from ONE torch.randn call on the current CUDA generator (the reference draws layer by layer at each
        # call: same distribution, different stream); options['bayes_z'] supplies it explicitly (tests, replays).
        stds = [p for l in info.linears() for p in (l.w_std, l.b_std)]
        n_theta = sum(p.numel() for p in params)
        if bayes_z is None:
            bayes_z = torch.randn(4 * n_steps, n_theta, dtype=torch.float32, device=y0.device)
        elif tuple(bayes_z.shape) != (4 * n_steps, n_theta) or bayes_z.device != y0.device or bayes_z.dtype != torch.float32:
            raise ValueError(f"bayes_z must be a float32 [{4 * n_steps}, {n_theta}] tensor on y0's device")
        need_tape = torch.is_grad_enabled() and (y0.requires_grad or any(p.requires_grad for p in (*params, *stds)))
        y_out, side = _Rk4SolveBayes.apply(y0, bayes_z, info, tables, n_steps, int(t.numel()), need_tape, len(params), *params, *stds)
    else:
        need_tape = torch.is_grad_enabled() and (y0.requires_grad or any(p.requires_grad for p in params))
        y_out, side = _Rk4Solve.apply(y0, info, tables, n_steps, int(t.numel()), need_tape, *params)
    if group is not None:
        from .dist import all_reduce_side
        side = all_reduce_side(side, None if group is True else group)
    install_side_channels(func, info, side)
    return y_out


def odeint_reparam(func, eps, std, mean, t, *, method="rk4", options=None):
    """``odeint(func, reparam(eps, std, mean, n_samples, batch) + 1e-5, t, method='rk4', options=...)`` with the producer of
    y0 fused into the solve (lib/models.py:16-24, lib/VAE.py:120,130,137 -- SURVEY 8(f1)): the ``[n_samples * batch, R, L]``
    initial state is never materialised by torch (it is output 0 of the solve), and ``loss.backward()`` returns the
    gradients of ``mean`` / ``std`` (/ ``eps``) straight from the adjoint kernel's dL/dz rows.

    eps ``[n_samples, batch, R, L-1]``, std / mean ``[batch, R, L-1]`` (the encoder's outputs).  Returns ``[len(t),
    n_samples * batch, R, L]`` with flat index ``s * batch + b`` like reparam's reshape."""
    if method != "rk4":
        raise NotImplementedError("the fused reparam prologue implements the reference's method, rk4")
    if eps.dim() != 4 or mean.shape != eps.shape[1:] or std.shape != eps.shape[1:]:
        raise ValueError("eps must be [n_samples, batch, R, L-1] and std / mean [batch, R, L-1]")
    info = describe_func(func)
    if info.bayes:
        raise NotImplementedError("Bayes_* funcs go through ude_b200.odeint (method='rk4')")
    if eps.shape[2] != info.n_regions or eps.shape[3] + 1 != info.latent_dim:
        raise ValueError(f"eps must be [n_samples, batch, {info.n_regions}, {info.latent_dim - 1}]")
    for x in (eps, std, mean):
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("the fused UDE solver runs on fp32 CUDA tensors only; there is no CPU fallback")
    options = {} if options is None else dict(options)
    group = options.pop("process_group", None)
    info.kernels = options.pop("kernels", os.environ.get("UDE_B200_KERNELS") or None)
    unknown = set(options) - {"step_size"}
    if unknown:
        raise ValueError(f"unexpected rk4 options {sorted(unknown)}")
    t_cpu = t.detach().to("cpu")
    if t.dim() != 1 or t.numel() < 2 or not bool((t_cpu[1:] > t_cpu[:-1]).all()):
        raise ValueError("t must be a one dimensional, strictly increasing tensor")
    (tables, n_steps) = _grid_on_device(t_cpu, options.get("step_size", None), eps.device)
    params = [p for l in info.linears() for p in (l.weight, l.bias)]
    need_tape = torch.is_grad_enabled() and (eps.requires_grad or std.requires_grad or mean.requires_grad or
                                             any(p.requires_grad for p in params))
    y_out, side = _Rk4SolveReparam.apply(eps, std, mean, info, tables, n_steps, int(t.numel()), need_tape, *params)
    if group is not None:
        from .dist import all_reduce_side
        side = all_reduce_side(side, None if group is True else group)
    install_side_channels(func, info, side)
    return y_out


def _decoder_linear(decoder) -> torch.nn.Linear:
    """The nn.Linear(3R, R * input_dim) behind the reference's Decoder (lib/models.py:38-42: Sequential(Flatten, Linear));
    an nn.Linear is taken as is."""
    if isinstance(decoder, torch.nn.Linear):
        return decoder
    seq = getattr(decoder, "decoder", None)
    lins = [m for m in (seq if seq is not None else decoder.modules()) if isinstance(m, torch.nn.Linear)]
    if len(lins) != 1:
        raise UnsupportedFunc("the fused decoder is ONE nn.Linear over latent[..., :3] (lib/models.py:26-51 with Fp=True)")
    if seq is not None and getattr(decoder, "latent_dim", 3) != 3:
        raise UnsupportedFunc("Decoder(Fp=False) reads all latent components; the fused epilogue covers the Fp=True decoder")
    return lins[0]


def odeint_decoded(func, decoder, y0, t, *, method="rk4", options=None, reparam=None):
    """The solve with its consumers fused into the output store.  For the reference's

        latent = odeint(self.ode, z, t, method='rk4', options=dict(step_size=h))          lib/VAE.py:137
        y      = self.dec(self.latent[..., :3])                                           lib/VAE.py:138
        reg    = train_functions.latent_init_loss(self.latent[..., :3])                   lib/VAE.py:186

    this returns ``(y, reg)`` -- ``y`` ``[len(t), B, R * input_dim]`` (what ``self.dec`` returns; the caller reshapes it to
    ``[batch, n_samples, T, R]`` as lib/VAE.py:138 does), ``reg`` a 0-dim tensor -- without ever writing the
    ``[len(t), B, R, L]`` trajectories: at L = 8 the kernels store 1/8 of the bytes (plus a byte of sign mask per region)
    and the adjoint reads ``dL/dy`` instead of a trajectory-sized cotangent.  Gradients reach the ODE func's parameters, the
    decoder's weight and bias, and y0 (or ``mean`` / ``std`` / ``eps`` with ``reparam=(eps, std, mean)``, y0 = None: the
    fused producer of SURVEY 8(f1))."""
    if method != "rk4":
        raise NotImplementedError("the fused decoder epilogue implements the reference's method, rk4")
    info = describe_func(func)
    lin = _decoder_linear(decoder)
    R, Ld = info.n_regions, info.latent_dim
    if lin.in_features != 3 * R or lin.bias is None:
        raise ValueError(f"the decoder must be nn.Linear({3 * R}, R * input_dim) with a bias")
    eps = std = mean = None
    if reparam is not None:
        if y0 is not None:
            raise ValueError("pass either y0 or reparam=(eps, std, mean)")
        eps, std, mean = reparam
        if eps.dim() != 4 or mean.shape != eps.shape[1:] or std.shape != eps.shape[1:] or eps.shape[2] != R or eps.shape[3] + 1 != Ld:
            raise ValueError(f"eps must be [n_samples, batch, {R}, {Ld - 1}] and std / mean [batch, {R}, {Ld - 1}]")
        ins = (eps, std, mean)
    else:
        if not isinstance(y0, torch.Tensor) or y0.dim() != 3 or y0.shape[1] != R or y0.shape[2] != Ld:
            raise ValueError(f"y0 must be [B, {R}, {Ld}]")
        ins = (y0,)
    for x in (*ins, lin.weight, lin.bias):
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("the fused UDE solver runs on fp32 CUDA tensors only; there is no CPU fallback")
    options = {} if options is None else dict(options)
    group = options.pop("process_group", None)
    info.kernels = options.pop("kernels", os.environ.get("UDE_B200_KERNELS") or None)
    bayes_z = options.pop("bayes_z", None)
    unknown = set(options) - {"step_size"}
    if unknown:
        raise ValueError(f"unexpected rk4 options {sorted(unknown)}")
    t_cpu = t.detach().to("cpu")
    if t.dim() != 1 or t.numel() < 2 or not bool((t_cpu[1:] > t_cpu[:-1]).all()):
        raise ValueError("t must be a one dimensional, strictly increasing tensor")
    (tables, n_steps) = _grid_on_device(t_cpu, options.get("step_size", None), lin.weight.device)
    params = [p for l in info.linears() for p in (l.weight, l.bias)]
    n_par = len(params)
    if info.bayes:      # weights re-drawn at every RHS evaluation (see odeint): params = means, then stds
        n_theta = sum(p.numel() for p in params)
        if bayes_z is None:
            bayes_z = torch.randn(4 * n_steps, n_theta, dtype=torch.float32, device=lin.weight.device)
        elif tuple(bayes_z.shape) != (4 * n_steps, n_theta) or bayes_z.device != lin.weight.device or bayes_z.dtype != torch.float32:
            raise ValueError(f"bayes_z must be a float32 [{4 * n_steps}, {n_theta}] tensor on the func's device")
        params = [*params, *[p for l in info.linears() for p in (l.w_std, l.b_std)]]
    elif bayes_z is not None:
        raise ValueError("options['bayes_z'] goes with a Bayes_* func")
    need_tape = torch.is_grad_enabled() and any(x.requires_grad for x in (*ins, lin.weight, lin.bias, *params))
    y_pred, side = _Rk4SolveDecoded.apply(y0, eps, std, mean, lin.weight, lin.bias, info, tables, n_steps, int(t.numel()),
                                          need_tape, bayes_z, n_par, *params)
    reg = side[7].to(torch.float32)          # this rank's trajectories only (a per-trajectory sum, like y_pred): a sharded trainer adds
    if group is not None:                    # its local terms and all-reduces gradients; posterior() / tracker need GLOBAL sums
        from .dist import all_reduce_side
        side = all_reduce_side(side, None if group is True else group)
    install_side_channels(func, info, side)
    return y_pred, reg


def flatten_group_parameters(funcs) -> torch.Tensor:
    """Re-home the weights of G same-architecture ODE funcs as rows of ONE ``[G, n_theta]`` leaf tensor (theta order per row) and
    return it: ``odeint_groups(funcs, ..., theta=flat)`` then hands the buffer to the kernels without gathering 6 G tensors, the
    gradient of the whole sweep arrives as ``flat.grad`` (one tensor: one optimizer over ``[flat]`` trains all G models -- Adam is
    element-wise), and the modules keep seeing the current values (their parameters become views of the rows)."""
    funcs = list(funcs)
    infos = [describe_func(f) for f in funcs]
    rows = [torch.cat([p.detach().reshape(-1) for l in i.linears() for p in (l.weight, l.bias)]) for i in infos]
    if len({r.numel() for r in rows}) != 1:
        raise ValueError("all funcs of a grouped solve must share one architecture")
    flat = torch.stack(rows).requires_grad_(True)
    for g, i in enumerate(infos):
        ps = [p for l in i.linears() for p in (l.weight, l.bias)]
        for p, v in zip(ps, split_like(flat.detach()[g], ps)):
            p.data = v
    return flat


def odeint_groups(funcs, y0, t, *, method="rk4", options=None, theta=None):
    """G independent UDE models integrated side by side (BASELINE configs[4]: the hyper-parameter sweeps of
    tuning/tune_node.py train every configuration with its own ``odeint`` call, :204,227 — here they share ONE).

    ``funcs``: G modules of identical architecture (own weights each); ``y0``: ``[G, B_g, R, L]``; returns
    ``[G, len(t), B_g, R, L]``, i.e. ``out[g] == odeint(funcs[g], y0[g], t, ...)``.  Gradients flow to every
    module's parameters and to ``y0``; each module receives its own side-channel entries (``params`` / ``tracker``).
    Fixed-grid rk4 only (the method of every live call site of the reference)."""
    if method != "rk4":
        raise NotImplementedError("grouped solves implement the reference's method, rk4")
    funcs = list(funcs)
    if not funcs:
        raise ValueError("funcs is empty")
    if not isinstance(y0, torch.Tensor) or y0.dim() != 4 or y0.shape[0] != len(funcs):
        raise ValueError(f"y0 must be [G={len(funcs)}, B_g, R, L]")
    if theta is not None:
        # the weights arrive as ONE [G, n_theta] tensor (flatten_group_parameters): funcs[0] describes the architecture of all
        infos = [_check_common(funcs[0], y0[0], t, method, None)]
        n_theta = sum(p.numel() for l in infos[0].linears() for p in (l.weight, l.bias))
        if tuple(theta.shape) != (len(funcs), n_theta) or theta.device != y0.device or theta.dtype != torch.float32:
            raise ValueError(f"theta must be a float32 [{len(funcs)}, {n_theta}] tensor on y0's device")
    else:
        infos = [_check_common(f, y0[g], t, method, None) for g, f in enumerate(funcs)]
    i0 = infos[0]
    if any(i.bayes for i in infos):
        raise NotImplementedError("Bayes_* funcs go through ude_b200.odeint (method='rk4')")
    sig0 = (i0.kind, i0.n_regions, i0.latent_dim, [l.weight.shape for l in i0.linears()], float(i0.fa_w),
            i0.rate_head_elu, i0.aug_head_elu)
    for i in infos[1:]:
        sig = (i.kind, i.n_regions, i.latent_dim, [l.weight.shape for l in i.linears()], float(i.fa_w), i.rate_head_elu,
               i.aug_head_elu)
        if sig != sig0:
            raise ValueError("all funcs of a grouped solve must share one architecture (and Fa_w)")
    options = {} if options is None else dict(options)
    i0.kernels = options.pop("kernels", os.environ.get("UDE_B200_KERNELS") or None)
    unknown = set(options) - {"step_size"}
    if unknown:
        raise ValueError(f"unexpected rk4 options {sorted(unknown)}")
    t_cpu = t.detach().to("cpu")
    if not bool((t_cpu[1:] > t_cpu[:-1]).all()):
        raise ValueError("t must be strictly increasing")
    (tables, n_steps) = _grid_on_device(t_cpu, options.get("step_size", None), y0.device)
    # one cat over all G x (W, b) tensors, one batched side-channel computation: the glue must not cost G x a dozen launches
    if theta is None:
        theta = torch.cat([p.reshape(-1) for i in infos for l in i.linears() for p in (l.weight, l.bias)]).view(len(funcs), -1)
    need_tape = torch.is_grad_enabled() and (theta.requires_grad or y0.requires_grad)
    y_out, side = _Rk4SolveGroups.apply(theta, y0, i0, tables, n_steps, int(t.numel()), need_tape)
    if i0.rate_layers:
        n = side[:, 5:6]
        mean = side[:, 0:2] / n
        half = torch.sqrt(torch.clamp((side[:, 2:4] - n * mean * mean) / (n - 1.0), min=0.0)) / math.sqrt(2.0)
        rows = torch.stack([mean - half, mean + half], dim=1).to(torch.float32)          # [G, 2, 2], see posterior_rows
        for g, f in enumerate(funcs):
            if isinstance(getattr(f, "params", None), list):
                f.params.append(rows[g])
    if i0.aug_layers:
        norms = torch.sqrt(side[:, 4]).to(torch.float32)
        for g, f in enumerate(funcs):
            if isinstance(getattr(f, "tracker", None), list):
                f.tracker.append(norms[g])
    return y_out
