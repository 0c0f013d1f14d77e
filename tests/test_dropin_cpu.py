"""Drop-in at the call surface, CPU part: the UNMODIFIED reference caller (lib/VAE.py) runs on top
of a module that is installed as ``torchdiffeq`` — here the oracle (the fused kernels need a GPU;
tests/test_dropin_gpu.py drives the same call idioms through the CUDA path).  Skipped where
/root/reference does not exist (the GPU box)."""
import os
import sys

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "lib")), reason="reference tree not present")


def test_unmodified_reference_vae_trains_through_an_installed_torchdiffeq():
    import oracle.odeint_ref as shim
    saved = sys.modules.get("torchdiffeq")
    sys.modules["torchdiffeq"] = shim                    # `from torchdiffeq import odeint` (lib/VAE.py:5)
    sys.path.insert(0, REF)
    try:
        import lib.models as models
        import lib.VAE as VAE
        torch.manual_seed(0)
        R, n_qs, L, window, gamma = 1, 9, 8, 8, 28
        vae = VAE.VAE(models.Encoder_Back_GRU, models.FaFp, models.Decoder, n_qs, L, n_regions=R,
                      ode_params=dict(net_sizes=[64, 64, 32], aug_net_sizes=[64, 64]),
                      enc_params=dict(q_sizes=[32, 16], ff_sizes=[16]))
        vae.setup_training(lr=1e-3)
        vae.skip_count = 0
        with torch.no_grad():
            vae.ode.aug_net[-1].weight.mul_(0.05)
            vae.ode.aug_net[-1].bias.mul_(0.05)
        batch = 4
        x = torch.rand(batch, window + 14, R * n_qs + R)
        y = torch.rand(batch, window + 1 + gamma, R) * 0.5
        t = torch.arange(window + gamma + 1, dtype=torch.float32) / 7
        eval_pts = list(range(0, 29, 7))
        losses = {'nll': True, 'mse': False, 'kl_z': True, 'kl_p': True, 'Fa_norm': 1e-1, 'reg_loss': True, 'anneal': True}
        data, names = vae.train_step(x, y, t, epoch=1, losses=losses, eval_pts=eval_pts, grad_lim=5000, n_samples=8)
        rec = dict(zip(names, data))
        assert all(k in rec for k in ("loss", "nll", "kl_params", "Fa_norm", "reg_loss", "grad_norm"))
        assert rec["grad_norm"] > 0 and rec["grad_norm"] == rec["grad_norm"]
        assert vae.latent.shape == (len(eval_pts), 8 * batch, R, L)
    finally:
        if saved is not None:
            sys.modules["torchdiffeq"] = saved
        else:
            sys.modules.pop("torchdiffeq", None)
        sys.path.remove(REF)


def test_package_exposes_the_torchdiffeq_surface():
    import ude_b200
    import inspect
    sig = inspect.signature(ude_b200.odeint)
    assert list(sig.parameters) == ["func", "y0", "t", "rtol", "atol", "method", "options", "event_fn"]
    assert sig.parameters["rtol"].default == 1e-7 and sig.parameters["atol"].default == 1e-9
    assert sig.parameters["method"].default is None          # None -> dopri5, as torchdiffeq
    adj = inspect.signature(ude_b200.odeint_adjoint)       # torchdiffeq's second entry point incl. its adjoint_* knobs
    assert list(adj.parameters)[:8] == list(sig.parameters)
    assert {"adjoint_rtol", "adjoint_atol", "adjoint_method", "adjoint_options", "adjoint_params"} <= set(adj.parameters)
