"""CPU tests of the host-side mirror of the reference's operator interface (no GPU)."""
import os

import pytest
import torch

import ude_b200
from oracle.ude_rhs import theta_from_module, ude_rhs
from tests.util import load_golden, module_from_spec


def test_state_dict_keys_follow_the_reference_layout():
    """Fp_net.{1,3,5,6}, net.* + aug_net.{1,3,4}, aug_net.{0,2,3}."""
    fp = ude_b200.Fp(1, 8, net_sizes=[64, 64, 32])
    assert list(fp.state_dict()) == [f"Fp_net.{i}.{p}" for i in (1, 3, 5, 6) for p in ("weight", "bias")]
    fafp = ude_b200.FaFp(1, 8, net_sizes=[64, 64, 32], aug_net_sizes=[64, 64])
    assert list(fafp.state_dict()) == [f"net.{i}.{p}" for i in (1, 3, 5, 6) for p in ("weight", "bias")] + \
        [f"aug_net.{i}.{p}" for i in (1, 3, 4) for p in ("weight", "bias")]
    fa = ude_b200.Fa(1, 8, aug_net_sizes=[64, 64])
    assert list(fa.state_dict()) == [f"aug_net.{i}.{p}" for i in (0, 2, 3) for p in ("weight", "bias")]
    # class defaults (lib/models.py:110,159,200)
    assert ude_b200.describe_func(ude_b200.Fp()).rate_layers[0].out_features == 20
    assert ude_b200.describe_func(ude_b200.FaFp()).aug_layers[0].out_features == 32
    # extra kwargs are swallowed like the reference's **kwargs (tune_encoders.py:155 passes nhidden=64)
    ude_b200.Fp(1, 8, nhidden=64, prior_std=0.05)


@pytest.mark.skipif(not os.path.isdir("/root/reference/lib"), reason="reference tree only exists in the build container")
def test_checkpoints_interchange_with_the_real_reference_modules():
    import sys
    sys.path.insert(0, "/root/reference")
    import lib.models as ref_models
    torch.manual_seed(0)
    for kind, kw in (("Fp", dict(net_sizes=[64, 64, 32])), ("Fa", dict(aug_net_sizes=[64, 64])),
                     ("FaFp", dict(net_sizes=[20, 20], aug_net_sizes=[32, 32]))):
        ref = getattr(ref_models, kind)(2, 8, **kw)
        ours = getattr(ude_b200, kind)(2, 8, **kw)
        ours.load_state_dict(ref.state_dict(), strict=True)
        x = torch.rand(5, 2, 8)
        assert torch.equal(ours(None, x), ref(None, x))
        # the reference modules are recognised as they are (duck typing)
        info = ude_b200.describe_func(ref)
        assert info.kind == kind and torch.equal(info.theta(), theta_from_module(ref)[1])


def test_eager_forward_equals_oracle_rhs_and_protocol():
    g = load_golden("rk4_fafp_r10")
    mod = module_from_spec(g["spec"], torch.from_numpy(g["theta"]), g["fa_w"])
    x = torch.from_numpy(g["y0"])
    f, p, fa = ude_rhs(g["spec"], torch.from_numpy(g["theta"]), x, g["fa_w"])
    assert torch.equal(mod(None, x), f)
    assert torch.equal(mod.params[0], p) and torch.equal(mod.tracker[0], fa)
    post = mod.posterior()
    assert mod.params == [] and post.mean.shape == (2,)
    mod.clear_tracking()
    assert mod.tracker == [] and mod.ode_type == "FaFp" and mod.uncertainty == "none"


def test_posterior_rows_reproduce_mean_and_unbiased_std():
    rows = torch.rand(1000, 2, dtype=torch.float64) + 0.1
    side = torch.tensor([rows[:, 0].sum(), rows[:, 1].sum(), (rows[:, 0] ** 2).sum(), (rows[:, 1] ** 2).sum(), 0.0,
                         1000.0, 0.0, 0.0], dtype=torch.float64)
    pr = ude_b200.posterior_rows(side)
    stacked = torch.stack([pr]).reshape(-1, 2)
    assert torch.allclose(stacked.mean(0).double(), rows.mean(0), rtol=1e-6)
    assert torch.allclose(stacked.std(0).double(), rows.std(0), rtol=1e-5)


def test_unsupported_inputs_fail_loudly():
    lin = torch.nn.Linear(3, 3)
    with pytest.raises(ude_b200.UnsupportedFunc):
        ude_b200.describe_func(lin)
    mod = ude_b200.Fp(1, 8)
    y0 = torch.rand(4, 1, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ude_b200.odeint(mod, y0, torch.arange(3.), method="rk4", options=dict(step_size=1.0))
    with pytest.raises(ValueError):
        ude_b200.odeint(mod, y0, torch.arange(3.), method="euler")
    bad = ude_b200.Fp(1, 8)
    bad.Fp_net[2] = torch.nn.ReLU()
    with pytest.raises(ude_b200.UnsupportedFunc):
        ude_b200.describe_func(bad)
