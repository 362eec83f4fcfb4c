# Energy of one molecule through the drop-in API (mirrors the reference's examples/single.py;
# with explicit charges: EEQ charges are outside the accelerated hot path).
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tad_dftd4_b200 as d4  # noqa: E402

dev = torch.device("cuda:0")
numbers = torch.tensor([6, 6, 6, 6, 7, 6, 16, 1, 1, 1, 1, 1], device=dev)  # C4NCS H5
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
q = torch.zeros(12, dtype=torch.float64, device=dev)  # e.g. from tad_multicharge.get_eeq_charges

param = d4.get_params(method="d4", functional="pbe0")
energy = d4.dftd4(numbers, positions, 0.0, param, q=q)
energy2 = d4.dispersion.DispD4().calculate(numbers, positions, 0.0, param, q=q)
assert torch.equal(energy, energy2)
print(energy)

cn, _, c6, alpha = d4.get_properties(numbers, positions, q=q)
print("CN   ", cn)
print("alpha", alpha)
