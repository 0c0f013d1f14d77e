"""Oracle restatement of ``torchdiffeq.odeint`` (TEST INFRASTRUCTURE ONLY).

The reference calls the third-party package ``torchdiffeq`` (un-vendored,
un-pinned; 0.2.3 at the reference's timestamps) at
/root/reference/lib/VAE.py:137, /root/reference/tuning/tune_node.py:204,227 and
seven more sites, always as ``odeint(func, y0, t, method='rk4',
options=dict(step_size=h))``.  The package is absent offline, so its published
algorithm is restated here (PARITY UNPINNED, see oracle/__init__.py):

* ``method='rk4'``  : FixedGridODESolver + ``rk4_alt_step_func`` (3/8 rule),
  grid from ``step_size``, linear interpolation of off-grid outputs
  (torchdiffeq/_impl/fixed_grid.py, solvers.py, rk_common.py).
* ``method='dopri5'`` (the default): RKAdaptiveStepsizeODESolver with the
  Dormand-Prince tableau, FSAL, fp64 time keeping, RMS error norm over the whole
  tensor, 4th-order dense output (torchdiffeq/_impl/rk_common.py, dopri5.py,
  interp.py, misc.py).
* ``options['replay_dt']`` (NOT in torchdiffeq; needs ``per_trajectory``): a ``[n, B]`` fp64 table of step sizes;
  trajectory ``b`` takes ``replay_dt[k, b]`` as its k-th step and accepts it -- the integrator and the dense output
  with the controller taken out, so that two implementations can be compared on the SAME step sequence.
* ``options['per_trajectory']=True`` (NOT in torchdiffeq): the same Dopri5
  controller applied to every trajectory (row of the leading batch axis) on its
  own, i.e. exactly what torchdiffeq does when called with B=1 — the semantics of
  the fused kernels (DESIGN.md "Dopri5 batch semantics").

Backward is plain autograd through these ops = the discrete adjoint, as in the
reference (/root/reference/lib/VAE.py:203).
"""
from __future__ import annotations

import math
from typing import Optional

import torch

_ONE_THIRD = 1.0 / 3.0
_TWO_THIRDS = 2.0 / 3.0


# ----------------------------------------------------------------------------
# fixed grid, rk4 (3/8 rule)
# ----------------------------------------------------------------------------
def rk4_grid(t: torch.Tensor, step_size) -> torch.Tensor:
    """FixedGridODESolver._grid_constructor_from_step_size: all arithmetic in t.dtype."""
    start_time = t[0]
    end_time = t[-1]
    niters = torch.ceil((end_time - start_time) / step_size + 1).item()
    t_infer = torch.arange(0, niters, dtype=t.dtype, device=t.device) * step_size + start_time
    t_infer[-1] = t[-1]
    return t_infer


def _rk4_alt_step(func, t0, dt, t1, y0, f0):
    k1 = f0
    k2 = func(t0 + dt * _ONE_THIRD, y0 + dt * k1 * _ONE_THIRD)
    k3 = func(t0 + dt * _TWO_THIRDS, y0 + dt * (k2 - k1 * _ONE_THIRD))
    k4 = func(t1, y0 + dt * (k1 - k2 + k3))
    return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125


def _linear_interp(t0, t1, y0, y1, t):
    if t == t0:
        return y0
    if t == t1:
        return y1
    slope = (t - t0) / (t1 - t0)
    return y0 + slope * (y1 - y0)


def _integrate_fixed_rk4(func, y0, t, step_size, grid_constructor=None):
    # torchdiffeq's FixedGridODESolver: an explicit grid_constructor(func, y0, t) wins over step_size (tests use it to run the
    # fp64 oracle on the very grid the fp32 arithmetic produces)
    if grid_constructor is not None:
        time_grid = grid_constructor(func, y0, t)
    else:
        time_grid = t if step_size is None else rk4_grid(t, step_size)
    assert time_grid[0] == t[0] and time_grid[-1] == t[-1]
    solution = [y0]
    j = 1
    for t0, t1 in zip(time_grid[:-1], time_grid[1:]):
        dt = t1 - t0
        f0 = func(t0, y0)
        y1 = y0 + _rk4_alt_step(func, t0, dt, t1, y0, f0)
        while j < len(t) and t1 >= t[j]:
            solution.append(_linear_interp(t0, t1, y0, y1, t[j]))
            j += 1
        y0 = y1
    assert j == len(t)
    return torch.stack(solution)


# ----------------------------------------------------------------------------
# dopri5
# ----------------------------------------------------------------------------
DOPRI5_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
DOPRI5_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
DOPRI5_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
DOPRI5_C_ERR = [
    35 / 384 - 1951 / 21600,
    0,
    500 / 1113 - 22642 / 50085,
    125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400,
    11 / 84 - 649 / 6300,
    -1. / 60.,
]
DOPRI5_C_MID = [
    6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2,
    -2691868925 / 45128329728 / 2, 187940372067 / 1594534317056 / 2,
    -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2,
]


def _side_lens(func):
    """Lengths of the reference's side-channel lists (lib/models.py:137,187,238,252)."""
    p = getattr(func, "params", None)
    f = getattr(func, "tracker", None)
    return (len(p) if isinstance(p, list) else 0, len(f) if isinstance(f, list) else 0)


def _side_pop(func, n_p0, n_f0):
    """Remove (and return) the side-channel rows appended since _side_lens()."""
    p = getattr(func, "params", None)
    f = getattr(func, "tracker", None)
    new_p, new_f = [], []
    if isinstance(p, list):
        new_p = p[n_p0:]
        del p[n_p0:]
    if isinstance(f, list):
        new_f = f[n_f0:]
        del f[n_f0:]
    return new_p, new_f


def _rms(x: torch.Tensor, per_traj: bool) -> torch.Tensor:
    """misc._rms_norm; per_traj -> one value per row of the leading axis."""
    if per_traj:
        return x.abs().pow(2).reshape(x.shape[0], -1).mean(1).sqrt()
    return x.abs().pow(2).mean().sqrt()


def _bc(v: torch.Tensor, like: torch.Tensor, per_traj: bool) -> torch.Tensor:
    """Broadcast a per-trajectory scalar against a state tensor."""
    if per_traj:
        return v.reshape(-1, *([1] * (like.dim() - 1)))
    return v


@torch.no_grad()
def _select_initial_step(func, t0, y0, order, rtol, atol, f0, per_traj):
    dtype = y0.dtype
    scale = atol + torch.abs(y0) * rtol
    d0 = _rms(y0 / scale, per_traj).abs()
    d1 = _rms(f0 / scale, per_traj).abs()
    small = (d0 < 1e-5) | (d1 < 1e-5)
    h0 = torch.where(small, torch.full_like(d0, 1e-6), 0.01 * d0 / d1).abs()
    y1 = y0 + _bc(h0, y0, per_traj) * f0
    n_p0, n_f0 = _side_lens(func)
    f1 = func(t0 + h0, y1)
    if per_traj:
        _side_pop(func, n_p0, n_f0)  # per-trajectory mode (not torchdiffeq's): the probe is not a side-channel row of the fused path
    d2 = torch.abs(_rms((f1 - f0) / scale, per_traj) / h0)
    tiny = (d1 <= 1e-15) & (d2 <= 1e-15)
    h1 = torch.where(tiny,
                     torch.max(torch.full_like(h0, 1e-6), h0 * 1e-3),
                     (0.01 / torch.max(d1, d2)) ** (1.0 / float(order + 1))).abs()
    return torch.min(100 * h0, h1).to(torch.float64)


@torch.no_grad()
def _optimal_step_size(last_step, error_ratio, safety=0.9, ifactor=10.0, dfactor=0.2, order=5):
    """misc._optimal_step_size, element-wise so it serves both batch modes (fp64)."""
    error_ratio = error_ratio.to(last_step.dtype)
    dfac = torch.where(error_ratio < 1, torch.ones_like(last_step), torch.full_like(last_step, dfactor))
    exponent = 1.0 / order
    safe_ratio = torch.where(error_ratio == 0, torch.ones_like(error_ratio), error_ratio)
    factor = torch.minimum(torch.full_like(last_step, ifactor),
                           torch.maximum(safety / safe_ratio ** exponent, dfac))
    factor = torch.where(error_ratio == 0, torch.full_like(last_step, ifactor), factor)
    return last_step * factor


def _interp_fit(y0, y1, y_mid, f0, f1, dt):
    a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
    b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
    c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
    d = dt * f0
    e = y0
    return [e, d, c, b, a]


def _interp_evaluate(coeffs, x):
    total = coeffs[0] + x * coeffs[1]
    x_power = x
    for c in coeffs[2:]:
        x_power = x_power * x
        total = total + x_power * c
    return total


def _weighted_sum(ks, ws, dt):
    """k.matmul(w * dt): sum_i k_i * (w_i * dt), accumulated left to right."""
    acc = None
    for k, w in zip(ks, ws):
        if w == 0:
            continue
        term = k * (dt * w)
        acc = term if acc is None else acc + term
    return acc


class Dopri5Stats:
    """Filled by odeint(..., options={'stats': Dopri5Stats()})."""
    def __init__(self):
        self.n_accept = None   # int (shared) or LongTensor[B] (per trajectory)
        self.n_reject = None
        self.dt0 = None        # initial step(s), fp64


def _integrate_dopri5(func, y0, t, rtol, atol, options):
    per_traj = bool(options.get("per_trajectory", False))
    first_step = options.get("first_step", None)
    max_num_steps = int(options.get("max_num_steps", 2 ** 31 - 1))
    safety = float(options.get("safety", 0.9))
    ifactor = float(options.get("ifactor", 10.0))
    dfactor = float(options.get("dfactor", 0.2))
    stats: Optional[Dopri5Stats] = options.get("stats", None)
    replay = options.get("replay_dt", None)
    if replay is not None:
        assert per_traj, "replay_dt needs per_trajectory=True"
        replay = replay.to(torch.float64)
    B = y0.shape[0]
    sdt = y0.dtype
    t64 = t.to(torch.float64)
    nb = B if per_traj else 1

    f0 = func(t[0], y0)
    if replay is not None:
        dt = replay[0].clone()
    elif first_step is None:
        dt = _select_initial_step(func, t[0], y0, 4, rtol, atol, f0, per_traj).reshape(nb)
    else:
        dt = torch.full((nb,), float(first_step), dtype=torch.float64)
    if stats is not None:
        stats.dt0 = dt.clone()
    t0c = torch.full((nb,), float(t64[0]), dtype=torch.float64)   # interval start of the interpolant
    t1c = t0c.clone()                                              # current time
    coeffs = [y0] * 5
    y_cur, f_cur = y0, f0
    n_acc = torch.zeros(nb, dtype=torch.long)
    n_rej = torch.zeros(nb, dtype=torch.long)

    solution = [y0]
    for j in range(1, len(t)):
        tj = t64[j]
        n_iter = 0
        while bool((tj > t1c).any()):
            assert n_iter < max_num_steps, "max_num_steps exceeded"
            n_iter += 1
            active = tj > t1c                                       # [nb]
            assert bool(((t1c + dt > t1c) | ~active).all()), "underflow in dt"
            dts = _bc(dt.to(sdt), y_cur, per_traj)                  # state-dtype step
            ks = [f_cur]
            n_p0, n_f0 = _side_lens(func)
            yi = None
            for alpha_i, beta_i in zip(DOPRI5_ALPHA, DOPRI5_BETA):
                yi = y_cur + _weighted_sum(ks, beta_i, dts)
                ks.append(func(t1c + alpha_i * dt, yi))
            # side channels: torchdiffeq (shared step) leaves EVERY evaluation in the reference's lists, rejected attempts
            # included; the per-trajectory mode of the fused path tracks ACCEPTED stages only: pull this
            # attempt's rows back out and re-append them below if accepted.
            side_p, side_f = _side_pop(func, n_p0, n_f0) if per_traj else ([], [])
            y1 = yi                                                 # c_sol == beta[-1] (FSAL)
            f1 = ks[-1]
            with torch.no_grad():
                err = _weighted_sum(ks, DOPRI5_C_ERR, dts)
                tol = atol + rtol * torch.max(y_cur.abs(), y1.abs())
                ratio = _rms(err / tol, per_traj).abs().reshape(nb)
                accept = (ratio <= 1) & active
                dt_next = _optimal_step_size(dt, ratio, safety, ifactor, dfactor, 5)
                if replay is not None:          # controller taken out: every step of the table is accepted as is
                    accept = active
                    k_next = torch.clamp(n_acc + accept.long(), max=replay.shape[0] - 1)
                    dt_next = replay.gather(0, k_next.reshape(1, -1)).reshape(nb)
            y_mid = y_cur + _weighted_sum(ks, DOPRI5_C_MID, dts)
            new_coeffs = _interp_fit(y_cur, y1, y_mid, ks[0], f1, dts)
            acc_b = _bc(accept, y_cur, per_traj)
            if per_traj:
                coeffs = [torch.where(acc_b, n, o) for n, o in zip(new_coeffs, coeffs)]
                y_cur = torch.where(acc_b, y1, y_cur)
                f_cur = torch.where(acc_b, f1, f_cur)
            elif bool(accept.all()):
                coeffs, y_cur, f_cur = new_coeffs, y1, f1
            if bool(accept.any()):
                for p in side_p:
                    func.params.append(p if not per_traj else p[accept])
                for fa in side_f:
                    func.tracker.append(fa if not per_traj else fa[accept])
            t0c = torch.where(accept, t1c, t0c)
            t1c = torch.where(accept, t1c + dt, t1c)
            dt = torch.where(active, dt_next, dt)
            n_acc += accept.long()
            n_rej += (active & ~accept).long()
        x = ((tj - t0c) / (t1c - t0c)).to(sdt)
        solution.append(_interp_evaluate(coeffs, _bc(x, y_cur, per_traj)))
    if stats is not None:
        stats.n_accept = n_acc if per_traj else int(n_acc[0])
        stats.n_reject = n_rej if per_traj else int(n_rej[0])
    return torch.stack(solution)


# ----------------------------------------------------------------------------
# public entry point (torchdiffeq/_impl/odeint.py:odeint)
# ----------------------------------------------------------------------------
def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None):
    if event_fn is not None:
        raise NotImplementedError("event handling is not on the reference's path")
    if not isinstance(y0, torch.Tensor):
        raise NotImplementedError("tuple states are not on the reference's path")
    if not torch.is_floating_point(y0):
        raise TypeError("`y0` must be a floating point Tensor")
    if not (isinstance(t, torch.Tensor) and t.dim() == 1 and torch.is_floating_point(t)):
        raise ValueError("t must be a one dimensional floating point tensor")
    if not bool((t[1:] > t[:-1]).all()):
        if bool((t[1:] < t[:-1]).all()):
            raise NotImplementedError("decreasing t is not on the reference's path")
        raise ValueError("t must be strictly increasing or decreasing")
    options = {} if options is None else dict(options)
    if method is None:
        method = "dopri5"
    if method == "rk4":
        return _integrate_fixed_rk4(func, y0, t, options.get("step_size", None), options.get("grid_constructor", None))
    if method == "dopri5":
        return _integrate_dopri5(func, y0, t, rtol, atol, options)
    raise ValueError(f"Invalid method \"{method}\"")
