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
from .funcs import (Bayes_Fa, Bayes_FaFp, Bayes_Fp, Decoder, Dense_Variational, Fa, FaFp, Fp, FuncInfo, UnsupportedFunc,  # noqa: F401
                    describe_func)
from . import dist, train  # noqa: F401
from .train import GraphedIteration, run_epoch  # noqa: F401
from .odeint import (flatten_group_parameters, flatten_parameters, odeint, odeint_decoded, odeint_groups, odeint_reparam, posterior_rows,  # noqa: F401
                     rk4_grid_tables)



def odeint_adjoint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None, adjoint_rtol=None,
                   adjoint_atol=None, adjoint_method=None, adjoint_options=None, adjoint_params=None):
    """torchdiffeq exports both entry points.  The fused backward IS the discrete adjoint of the forward solve, so the
    ``adjoint_*`` knobs of the continuous-adjoint solver have nothing to configure here: accepted and ignored."""
    return odeint(func, y0, t, rtol=rtol, atol=atol, method=method, options=options, event_fn=event_fn)


__all__ = ["odeint", "odeint_groups", "odeint_adjoint", "Fp", "Fa", "FaFp", "describe_func", "build", "lib", "UdeError",
           "UnsupportedFunc", "rk4_grid_tables", "posterior_rows", "flatten_parameters", "flatten_group_parameters", "odeint_reparam", "odeint_decoded", "Decoder", "Bayes_Fp", "Bayes_Fa", "Bayes_FaFp", "Dense_Variational", "GraphedIteration", "run_epoch"]
__version__ = "0.1.0"
