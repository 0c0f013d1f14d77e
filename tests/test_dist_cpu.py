"""Multi-rank host logic on CPU (gloo, world_size 2): sharding along the `batch` sub-index, the
differentiable side-sum all-reduce and the gradient all-reduce.  The per-rank 'solver' here is the
oracle (test infrastructure) — the point is that shard -> solve -> all-reduce reproduces the
single-rank result, which is exactly how bench.py / a training loop drive the fused kernels."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _loss(y, func, n_rows_total, side_reduce):
    """Data term (sum over trajectories) + posterior/Fa-norm terms formed from the GLOBAL side sums."""
    from ude_b200.odeint import posterior_rows
    side = side_reduce(func.side_sums())
    rows = posterior_rows(torch.cat([side, torch.zeros(2, dtype=side.dtype)]))
    mean, std = rows.mean(0), rows.std(0)
    return (y[..., 1] ** 2).sum() + 3.0 * mean.sum() + 2.0 * std.sum() + 0.1 * torch.sqrt(side[4])


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ude_b200
    from oracle import odeint_ref
    from oracle.ude_rhs import OracleFunc, UdeSpec, init_theta, synth_y0
    torch.set_num_threads(1)
    spec = UdeSpec("FaFp", 1, 8, (16, 16), (16,))
    theta = init_theta(spec, seed=0)
    n_samples, batch = 4, 5
    y0 = synth_y0(n_samples * batch, 1, 8, seed=1)
    t = torch.arange(4.)

    # single-rank reference (computed redundantly on every rank)
    f_all = OracleFunc(spec, theta)
    y_all = odeint_ref.odeint(f_all, y0, t, method="rk4", options=dict(step_size=1.0))
    _loss(y_all, f_all, None, lambda s: s[:6]).backward()

    # sharded: local solve, global side sums, all-reduced grads
    y0_loc = ude_b200.dist.shard_flat(y0, n_samples, batch, world, rank)
    f_loc = OracleFunc(spec, theta)
    y_loc = odeint_ref.odeint(f_loc, y0_loc, t, method="rk4", options=dict(step_size=1.0))
    _loss(y_loc, f_loc, None, lambda s: ude_b200.dist.all_reduce_side(s)[:6]).backward()
    # the data term is a SUM over trajectories and the side terms are global: after the all-reduce of
    # the gradients the side-term contribution has been added `world` times to nothing — each rank only
    # differentiated its own rows — so the sum over ranks is the single-rank gradient.
    ude_b200.dist.all_reduce_grads(f_loc.parameters())

    # (different batch sizes take different BLAS paths: equal to fp32 round-off, not bit for bit)
    ok_shard = torch.allclose(ude_b200.dist.shard_time_major(y_all.detach(), n_samples, batch, world, rank), y_loc.detach(), rtol=1e-5, atol=1e-6)
    y_back = ude_b200.dist.gather_time_major(y_loc.detach(), n_samples, batch)
    ok_gather = torch.allclose(y_back, y_all.detach(), rtol=1e-5, atol=1e-6) and torch.equal(
        ude_b200.dist.shard_time_major(y_back, n_samples, batch, world, rank), y_loc.detach())
    err = float((f_loc.theta.grad - f_all.theta.grad).abs().max() / f_all.theta.grad.abs().max())
    lo, hi = ude_b200.dist.batch_bounds(batch, world, rank)
    ret[rank] = (ok_shard, ok_gather, err, (lo, hi))
    dist.destroy_process_group()


def test_two_rank_sharded_solve_matches_single_rank():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        ok_shard, ok_gather, err, bounds = ret[rank]
        assert ok_shard, "local solve differs from the slice of the global solve"
        assert ok_gather, "gather_time_major is not the inverse of shard_time_major"
        assert err < 1e-5, f"all-reduced gradient differs from the single-rank gradient: {err}"
    assert ret[0][3] == (0, 3) and ret[1][3] == (3, 5)


def test_batch_bounds_cover_everything():
    import ude_b200
    for batch in (1, 7, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [ude_b200.dist.batch_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
