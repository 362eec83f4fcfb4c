"""Reproduces (and checks the fix for) the exit-time abort analysed in bench.py:leave().

    python tools/exit_repro.py race   # backward through a Python autograd.Function, then exit at once
    python tools/exit_repro.py idle   # same, but the GIL is released for 0.25 s before the exit

The exit code is the evidence (134 = std::terminate from an autograd worker thread)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench_inputs  # noqa: E402
import tad_dftd4_b200 as d4  # noqa: E402

mode = sys.argv[1]
numbers, positions, q = (torch.as_tensor(t) for t in bench_inputs.organic_batch([100] * 64, seed=3))
dev = torch.device("cuda:0")
numbers, positions, q = numbers.to(dev), positions.to(dev).double(), q.to(dev).double()
param = dict(s6=1.0, s8=1.2, s9=1.0, a1=0.4, a2=5.0)
for _ in range(3):
    pos = positions.detach().requires_grad_(True)
    e = d4.dftd4(numbers, pos, 0.0, param, q=q)
    (g,) = torch.autograd.grad(e.sum(), pos)
if mode == "idle":
    time.sleep(0.25)
print("ok", float(g.abs().max()))
