# Analytic forces through torch.autograd (mirrors the reference's examples/forces.py).
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
q = torch.zeros(12, dtype=torch.float64, device=dev)
param = d4.get_params(method="d4", functional="tpssh")

pos = positions.clone().requires_grad_(True)
energy = d4.dftd4(numbers, pos, 0.0, param, q=q)
(grad,) = torch.autograd.grad(energy.sum(), pos)

# central differences
num = torch.zeros_like(positions)
step = 1e-4
for i in range(numbers.shape[-1]):
    for j in range(3):
        p = positions.clone()
        p[i, j] += step
        e1 = d4.dftd4(numbers, p, 0.0, param, q=q).sum()
        p[i, j] -= 2 * step
        e2 = d4.dftd4(numbers, p, 0.0, param, q=q).sum()
        num[i, j] = (e1 - e2) / (2 * step)
assert torch.allclose(grad, num, atol=1e-8), (grad - num).abs().max()
print(grad)
