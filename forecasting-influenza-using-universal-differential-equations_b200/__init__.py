"""B200-native fused UDE solve: drop-in ``odeint`` + the reference's ODE-func modules.

Usage (the reference's own idiom, /root/reference/lib/VAE.py:5,137)::

    from ude_b200 import odeint, FaFp            # instead of torchdiffeq / lib.models
    latent = odeint(ode, z, t, method='rk4', options=dict(step_size=t[1] - t[0]))

or, without touching the reference's sources::

    import sys, ude_b200; sys.modules['torchdiffeq'] = ude_b200

The directory is named after the reference repository; ``ude_b200`` (repo root) is the
importable alias of this package.
"""
from ._lib import UdeError, build, lib  # noqa: F401
from .funcs import Fa, FaFp, Fp, FuncInfo, UnsupportedFunc, describe_func  # noqa: F401
from . import dist  # noqa: F401
from .odeint import odeint, odeint_groups, posterior_rows, rk4_grid_tables  # noqa: F401

odeint_adjoint = odeint   # torchdiffeq exports both; the fused backward IS the discrete adjoint

__all__ = ["odeint", "odeint_groups", "odeint_adjoint", "Fp", "Fa", "FaFp", "describe_func", "build", "lib", "UdeError",
           "UnsupportedFunc", "rk4_grid_tables", "posterior_rows"]
__version__ = "0.1.0"
