"""CUDA-graphed training iteration and an on-GPU epoch driver around the fused solve.

The reference's hot loop (/root/reference/lib/VAE.py:200-223 ``train_step``, :248-291 ``train``; curriculum over the
horizon /root/reference/run_ode.py:147-164) issues, per iteration, the encoder, ``odeint``, the decoder, a dozen loss
ops, ``loss.backward()``, a gradient-norm gate and an Adam step, and reads the loss back (``.item()``) every time.  With
the solve fused into two kernels the iteration is launch- and host-bound at the reference's batch sizes (32 x 64 = 2048
trajectories), so this module

* captures ONE iteration -- forward through ``ude_b200.odeint`` (any rk4 call of this package: its launches go to the
  current stream, its scratch comes from torch's graph-aware allocator), the user's loss, ``backward()`` and, when the
  optimizer is capturable, the step with the reference's gradient-norm gate done ON the device -- into a CUDA graph per
  (input shapes, time grid), and replays it;
* drives an epoch without a host round trip per iteration: losses and gate decisions accumulate in device tensors and
  are read once at the end.

No CPU path, no fallback: a capture that fails raises.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Optional, Sequence, Tuple

import torch


def _key(inputs: Sequence[torch.Tensor], tag) -> tuple:
    return (tag,) + tuple((tuple(x.shape), x.dtype, x.device.index) for x in inputs)


class GraphedIteration:
    """``loss = it(*inputs)`` runs one captured training iteration.

    step_fn(*inputs) -> scalar loss tensor, built with this package's ``odeint`` (method='rk4').
    params: the tensors ``loss.backward()`` fills (``.grad`` buffers are kept static between replays).
    optimizer: optional; captured too when it is CUDA-graph capturable (``torch.optim.Adam(..., capturable=True)``).
    grad_lim: the reference's gate (lib/VAE.py:205-212): the step is applied only while the global gradient norm stays
      below it -- here as a device-side blend, so a gated iteration leaves the parameters AND the optimizer state of the
      previous iteration in place without a host decision.
    tag: anything hashable that changes the captured work without changing input shapes (e.g. the time grid's bytes).
    after_backward: optional ``fn(loss) -> tensor | None`` run INSIDE the captured iteration once the gradients sit in
      ``p.grad`` and before the optimizer: the place for the gradient all-reduce of a sharded run (NCCL collectives are
      capturable); what it returns is kept as ``it.last_extra`` (a static tensor: e.g. the flat reduced gradient to read back).
    device: where the captured work runs when the inputs are HOST tensors (pinned memory): every call then starts with the
      asynchronous host -> device copy of the inputs into the graph's static buffers.
    """

    def __init__(self, step_fn: Callable[..., torch.Tensor], params: Iterable[torch.nn.Parameter],
                 optimizer: Optional[torch.optim.Optimizer] = None, grad_lim: Optional[float] = None, warmup: int = 3,
                 after_backward: Optional[Callable[[torch.Tensor], Optional[torch.Tensor]]] = None, device=None):
        self.step_fn = step_fn
        self.params = [p for p in params]
        self.optimizer = optimizer
        self.grad_lim = grad_lim
        self.warmup = warmup
        self.after_backward = after_backward
        self.device = torch.device(device) if device is not None else None
        self.last_extra = None
        self._graphs: Dict[tuple, Tuple[torch.cuda.CUDAGraph, list, torch.Tensor, torch.Tensor, torch.Tensor]] = {}

    # ---- one eager iteration (also the warm-up and the body that is captured)
    def _iteration(self, inputs):
        loss = self.step_fn(*inputs)
        # torch.autograd.grad + explicit writes into the (static) .grad buffers instead of loss.backward(): no AccumulateGrad
        # nodes take part, so parameters that were trained eagerly before (their accumulator nodes remember the stream they
        # were created on) can be captured all the same
        got = torch.autograd.grad(loss, self.params, allow_unused=True)
        with torch.no_grad():
            for p, g in zip(self.params, got):
                if g is None:
                    if p.grad is not None:
                        p.grad.zero_()
                elif p.grad is None:
                    p.grad = g.detach().clone()
                else:
                    p.grad.copy_(g)
        extra = self.after_backward(loss.detach()) if self.after_backward is not None else None
        grads = [p.grad for p in self.params if p.grad is not None]
        norm = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads])) if grads else loss.new_zeros(())
        ok = torch.ones((), device=loss.device)
        if self.optimizer is not None:
            if self.grad_lim is None:
                self.optimizer.step()
            else:
                # device-side gate: take the step, then blend it back when the norm is over the limit
                ok = (norm < self.grad_lim).to(loss.dtype)
                before = [p.detach().clone() for p in self.params]
                state_before = [{k: v.clone() for k, v in self.optimizer.state[p].items() if torch.is_tensor(v)}
                                for p in self.params if p in self.optimizer.state]
                self.optimizer.step()
                with torch.no_grad():
                    for p, b in zip(self.params, before):
                        p.copy_(torch.lerp(b, p, ok))
                    i = 0
                    for p in self.params:
                        if p in self.optimizer.state and i < len(state_before):
                            for k, v in state_before[i].items():
                                cur = self.optimizer.state[p][k]
                                cur.copy_(torch.lerp(v.to(cur.dtype), cur, ok.to(cur.dtype)))
                            i += 1
        return loss.detach(), norm.detach(), ok, extra

    def __call__(self, *inputs: torch.Tensor, tag=None):
        key = _key(inputs, tag)
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._capture(inputs, key)
        graph, static_in, loss, norm, ok, extra = entry
        with torch.no_grad():       # static inputs may be leaves that require grad (dL/dy0 wanted)
            for s, x in zip(static_in, inputs):
                if s.data_ptr() != x.data_ptr():
                    s.copy_(x, non_blocking=True)          # host inputs: the H2D copy of this iteration, on the replay's stream
        graph.replay()
        self.last_norm, self.last_ok, self.last_extra = norm, ok, extra
        return loss

    def _capture(self, inputs, key):
        # Autograd graphs of earlier EAGER iterations may linger in reference cycles; they keep the parameters' gradient
        # accumulator nodes alive, and those remember the (legacy default) stream they were created on -- which the capture
        # stream must not synchronise with.  Collect them so that the warm-up below re-creates the nodes on its own stream.
        import gc
        gc.collect()
        dev = self.device if self.device is not None else inputs[0].device
        static_in = [x.detach().to(dev, copy=True).requires_grad_(x.requires_grad) for x in inputs]
        # optimizer state must exist before capture (lazy init), grads must be allocated: warm up on a side stream
        snapshot = [p.detach().clone() for p in self.params]
        opt_state = {}
        if self.optimizer is not None:      # a capture in mid-training (new horizon of the curriculum) keeps the optimizer's state
            opt_state = {p: {k: v.clone() for k, v in st.items() if torch.is_tensor(v)} for p, st in self.optimizer.state.items()}
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(self.warmup):
                self._iteration(static_in)
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        # the warm-up iterations moved the parameters and the optimizer state: put both back (state that did not exist
        # before the warm-up starts from zero, as a fresh optimizer's would)
        with torch.no_grad():
            for p, b in zip(self.params, snapshot):
                p.copy_(b)
            if self.optimizer is not None:
                for p, st in self.optimizer.state.items():
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            old = opt_state.get(p, {}).get(k)
                            v.copy_(old) if old is not None else v.zero_()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss, norm, ok, extra = self._iteration(static_in)
        entry = (graph, static_in, loss, norm, ok, extra)
        self._graphs[key] = entry
        return entry


def run_epoch(it: GraphedIteration, batches: Iterable[Sequence[torch.Tensor]], tag=None) -> dict:
    """One epoch through a GraphedIteration without a host round trip per iteration (lib/VAE.py:256-262 reads the loss
    back after every step).  Returns {'loss': mean loss, 'steps_taken': ..., 'iterations': ...} read ONCE at the end."""
    total = None
    taken = None
    n = 0
    for batch in batches:
        loss = it(*batch, tag=tag)
        total = loss.clone() if total is None else total + loss
        taken = it.last_ok.clone() if taken is None else taken + it.last_ok
        n += 1
    if n == 0:
        return {"loss": float("nan"), "steps_taken": 0, "iterations": 0}
    out = torch.stack([total / n, taken]).cpu()
    return {"loss": float(out[0]), "steps_taken": int(round(float(out[1]))), "iterations": n}
