"""Small EEQ calls (both CTA sizes, forward + VJP, padding, float32 I/O) for compute-sanitizer memcheck / racecheck."""
import pathlib
import sys

root = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(root))
import torch

import bench_inputs
from tad_dftd4_b200 import eeq

dev = torch.device("cuda:0")
for sizes in ([1, 2, 5, 31, 32, 33, 40, 63, 64], [3, 65, 100, 128, 97], [160, 1, 150]):
    numbers, positions, _ = bench_inputs.organic_batch(sizes, seed=4)
    for dtype in (torch.float64, torch.float32):
        n, p = numbers.to(dev), positions.to(dev, dtype)
        charge = torch.zeros(len(sizes), dtype=dtype, device=dev)
        eng = eeq._EeqEngine.get(dev)
        q = eng.charges(n, p, charge, 25.0)
        g = eng.vjp(n, p, 25.0, q, torch.ones_like(q))
        q, factor = eng.charges(n, p, charge, 25.0, keep_factor=True)
        g2 = eng.vjp(n, p, 25.0, q, torch.ones_like(q), factor)
        assert float((g - g2).abs().max()) < 1e-4
        print(sizes, dtype, float(q.sum().abs()), float(g.abs().max()))
torch.cuda.synchronize()
