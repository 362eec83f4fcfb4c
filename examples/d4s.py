# D4S model: element-pair specific Gaussian weighting (mirrors the reference's examples/d4s.py).
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tad_dftd4_b200 as d4  # noqa: E402

dev = torch.device("cuda:0")
numbers = torch.tensor([14, 1, 1, 1, 1], device=dev)  # SiH4
a = 1.61768389755830
positions = torch.tensor([[0, 0, 0], [a, a, -a], [-a, -a, -a], [a, -a, a], [-a, a, a]], dtype=torch.float64, device=dev)
q = torch.tensor([-8.412842390895063e-02] + [2.103210597723753e-02] * 4, dtype=torch.float64, device=dev)
param = d4.get_params(method="d4", functional="tpssh")
print("D4 ", d4.dftd4(numbers, positions, 0.0, param, q=q))
print("D4S", d4.dftd4(numbers, positions, 0.0, param, q=q, model=d4.model.D4SModel(numbers)))
print("D4S", d4.dftd4(numbers, positions, 0.0, param, q=q, model="d4s"))
