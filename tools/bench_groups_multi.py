"""BASELINE configs[4] at full size: 512 independent Fp(R=1, L=8, [32,32]) models x 2048 trajectories, 64 per GPU on 8 GPUs
(torchrun --nproc-per-node N), one grouped forward + one grouped adjoint launch per rank and iteration, NO collective
(tuning/tune_node.py trains every configuration on its own).  Rank 0 prints one JSON line (max over ranks)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import ude_b200
from oracle.ude_rhs import UdeSpec, init_theta, synth_y0
from tests.util import module_from_spec

world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
G_total, Bg = 512, 2048
G = G_total // 8 if world == 1 else G_total // world
spec = UdeSpec("Fp", 1, 8, (32, 32), ())
mods = [module_from_spec(spec, init_theta(spec, rank * G + g), 1.0, dev) for g in range(G)]
y0 = synth_y0(G * Bg, 1, 8, seed=3 + rank).reshape(G, Bg, 1, 8).to(dev)
t = torch.linspace(1, 56, 56) / 7

flat = ude_b200.flatten_group_parameters(mods)       # the weights of this rank's models as ONE [G, n_theta] leaf (one optimizer trains them all)

def step():
    y0g = y0.clone().requires_grad_(True)
    flat.grad = None
    y = ude_b200.odeint_groups(mods, y0g, t, method="rk4", options=dict(step_size=1.0), theta=flat)
    (y[..., 1] ** 2).mean().backward()

for _ in range(3):
    step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 5], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(ms)
    print(json.dumps({"config": f"configs[4]: {G * world} weight-set groups x Fp(R=1,L=8,[32,32]) x {Bg} trajectories over {world} GPU(s), "
                                "RK4 h=1 t=[1/7..8] (8 grid steps, 56 outputs), ONE grouped fwd + ONE grouped adjoint launch per rank (tcgen05 engine), weights as one [G, n_theta] leaf, no collective",
                      "n_gpus": world, "groups_per_gpu": G, "ms_per_iter": ms, "trajectory_steps_per_s": world * G * Bg * 8 / (ms * 1e-3)}))
if world > 1:
    dist.destroy_process_group()
