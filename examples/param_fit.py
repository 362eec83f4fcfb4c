# Differentiable damping parameters: gradients of the dispersion energy with respect to
# s6, s8, s9, s10, a1, a2, alp (the reference differentiates its dense tape,
# test/test_grad/test_param.py; here: d4b200_param_vjp_* behind the same autograd interface),
# used to fit s8/a1/a2 of a functional to a target energy with a torch optimiser.
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tad_dftd4_b200 as d4  # noqa: E402

dev = torch.device("cuda:0")
numbers = torch.tensor([6, 6, 6, 6, 7, 6, 16, 1, 1, 1, 1, 1], device=dev)
positions = torch.tensor(
    [
        [-2.56745685564671, -0.02509985979910, 0.0], [-1.39177582455797, +2.27696188880014, 0.0],
        [+1.27784995624894, +2.45107479759386, 0.0], [+2.62801937615793, +0.25927727028120, 0.0],
        [+1.41097033661123, -1.99890996077412, 0.0], [-1.17186102298849, -2.34220576284180, 0.0],
        [-2.39505990368378, -5.22635838332362, 0.0], [+2.41961980455457, -3.62158019253045, 0.0],
        [-2.51744374846065, +3.98181713686746, 0.0], [+2.24269048384775, +4.24389473203647, 0.0],
        [+4.66488984573956, +0.17907568006409, 0.0], [-4.60044244782237, -0.17794734637413, 0.0],
    ],
    dtype=torch.float64, device=dev,
)  # fmt: skip

# TPSS0-D4-ATM (test/test_grad/test_param.py:54-62 of the reference), every parameter differentiable
values = dict(s6=1.0, s8=0.78981345, s9=1.0, s10=0.0, a1=0.49484001, a2=5.73083694, alp=16.0)
param = d4.Param(**{k: torch.tensor(v, dtype=torch.float64, device=dev, requires_grad=True) for k, v in values.items()})
energy = d4.dftd4(numbers, positions, torch.tensor(0.0, device=dev, dtype=torch.float64), param)  # q=None: EEQ charges
grads = torch.autograd.grad(energy.sum(), list(param.values()))
print(f"E = {energy.sum().item():.10e} Eh")
for k, g in zip(param, grads):
    print(f"  dE/d{k:<3s} = {g.item():+.10e}")

# fit s8, a1, a2 so that the dispersion energy of this molecule becomes 5 % more negative
target = 1.05 * energy.sum().item()
fit = {k: torch.tensor(values[k], dtype=torch.float64, device=dev, requires_grad=True) for k in ("s8", "a1", "a2")}
opt = torch.optim.Adam(list(fit.values()), lr=5e-3)
for it in range(60):
    opt.zero_grad()
    loss = (d4.dftd4(numbers, positions, 0.0, d4.Param(**fit)).sum() - target) ** 2
    loss.backward()
    opt.step()
print("fitted:", {k: round(v.item(), 6) for k, v in fit.items()}, f"loss = {loss.item():.3e}")
