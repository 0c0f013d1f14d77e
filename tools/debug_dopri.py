import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ude_b200
from oracle import odeint_ref
from oracle.ude_rhs import OracleFunc, UdeSpec, init_theta, synth_y0
from tests.util import module_from_spec, rel_err, flat_grad

spec = UdeSpec("Fp", 1, 8, (16, 16), ())
theta = init_theta(spec, seed=3)
y0 = synth_y0(4, 1, 8, seed=5)
for tname, t in [("t2", torch.tensor([0., 8.])), ("t9", torch.arange(9.)), ("tmid", torch.tensor([0., 0.01, 0.02, 8.0]))]:
    for lname in ["all", "last", "first", "comp0_last"]:
        def loss_of(y):
            if lname == "all": return (y ** 2).sum()
            if lname == "last": return (y[-1] ** 2).sum()
            if lname == "first": return (y[1] ** 2).sum()
            return y[-1, :, :, 0].sum()
        f = OracleFunc(spec, theta)
        y0r = y0.clone().requires_grad_(True)
        st = odeint_ref.Dopri5Stats()
        yr = odeint_ref.odeint(f, y0r, t, rtol=1e-5, atol=1e-9, options=dict(per_trajectory=True, stats=st))
        loss_of(yr).backward()
        mod = module_from_spec(spec, theta, 1.0, "cuda")
        y0g = y0.cuda().requires_grad_(True)
        y, stats = ude_b200.odeint(mod, y0g, t, rtol=1e-5, atol=1e-9, options=dict(return_stats=True))
        loss_of(y).backward()
        print(tname, lname, "acc", stats.n_accept.tolist(), st.n_accept.tolist(), "rej", stats.n_reject.tolist(), st.n_reject.tolist(),
              "y %.2e dy0 %.2e dth %.2e" % (rel_err(y, yr), rel_err(y0g.grad, y0r.grad), rel_err(flat_grad(mod), f.theta.grad)))
