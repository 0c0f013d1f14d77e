import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ude_b200
from oracle import odeint_ref
from oracle.ude_rhs import OracleFunc
from tests.util import module_from_spec, rel_err, flat_grad, load_golden

# deterministic controller: tolerance so loose that every step is accepted and grows by ifactor exactly
for name in ["dopri5_fp_r3", "dopri5_fafp_r1"]:
    g = load_golden(name)
    spec = g["spec"]; theta = torch.from_numpy(g["theta"]); y0 = torch.from_numpy(g["y0"]); t = torch.from_numpy(g["t"])
    G = torch.from_numpy(g["G"])
    kw = dict(rtol=1e-1, atol=1e-1)
    for lname in ["yG", "side_b", "side_b2", "side_f2"]:
        def loss_of(y, side):
            if lname == "yG": return (y.double() * G.to(y.device).double()).sum()
            if lname == "side_b": return side[0] + 0.5 * side[1]
            if lname == "side_b2": return side[2] + side[3]
            return side[4]
        if lname == "side_f2" and not spec.has_aug: continue
        if lname in ("side_b", "side_b2") and not spec.has_rate: continue
        f = OracleFunc(spec, theta, g["fa_w"])
        y0r = y0.clone().requires_grad_(True)
        st = odeint_ref.Dopri5Stats()
        yr = odeint_ref.odeint(f, y0r, t, options=dict(per_trajectory=True, stats=st, first_step=0.01), **kw)
        loss_of(yr, f.side_sums()).backward()
        mod = module_from_spec(spec, theta, g["fa_w"], "cuda")
        y0g = y0.cuda().requires_grad_(True)
        mod.clear_tracking()
        y, stats = ude_b200.odeint(mod, y0g, t, options=dict(return_stats=True, first_step=0.01), **kw)
        n = float(f.side_sums()[5])
        if spec.has_rate:
            rows = torch.stack(mod.params).reshape(-1, 2).double(); mean, var = rows.mean(0), rows.var(0)
            s1 = mean * n; s2 = var * (n - 1) + n * mean * mean
        else:
            s1 = s2 = torch.zeros(2, dtype=torch.float64, device="cuda")
        s4 = torch.stack(mod.tracker).double().pow(2).sum() if spec.has_aug else torch.zeros((), dtype=torch.float64, device="cuda")
        side = torch.cat([s1, s2, s4.reshape(1)])
        loss_of(y, side).backward()
        print(name, lname, "acc", stats.n_accept.tolist(), st.n_accept.tolist(), "rej", stats.n_reject.tolist(), st.n_reject.tolist(),
              "y %.2e side %.2e dy0 %.2e dth %.2e" % (rel_err(y, yr), rel_err(side, f.side_sums()[:5]), rel_err(y0g.grad, y0r.grad), rel_err(flat_grad(mod), f.theta.grad)))
