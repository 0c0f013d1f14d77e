"""CPU oracle for the fused UDE solve (TEST INFRASTRUCTURE ONLY).

This package restates, in plain PyTorch on the CPU, the reference's hot path:
the SIR/MLP right-hand sides of ``/root/reference/lib/models.py:109-265`` and
the third-party ``torchdiffeq.odeint`` solver the reference calls at
``/root/reference/lib/VAE.py:137``.

PARITY UNPINNED: the reference ships no tests, golden vectors or seeds
(SURVEY.md section 4), and ``torchdiffeq`` (un-vendored, un-pinned; PyPI 0.2.3
was current at the reference's timestamps) is neither installed nor
installable in the build container.  The solver here is restated from the
published algorithm of torchdiffeq 0.2.x.  The RHS restatement IS pinned: it is
checked bit-for-bit against the real ``lib.models`` classes imported from
``/root/reference`` (``tests/golden/make_golden.py`` generated the committed
fixtures with those classes).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package never
does.
"""
