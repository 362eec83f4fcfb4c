"""Times the EEQ kernels alone (charges, and the position VJP) on the C2 and C3 batch shapes with CUDA events.

    [D4B200_LIBRARY=build_ab/x.so] python tools/eeq_time.py
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench_inputs  # noqa: E402
from tad_dftd4_b200 import eeq  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


rng = np.random.default_rng(0)
for name, sizes in (("C2 4096 x 20-60", rng.integers(20, 61, 4096)), ("C3 1024 x 100", [100] * 1024),
                    ("512 x 160", [160] * 512)):
    numbers, positions, _ = bench_inputs.organic_batch_parallel(sizes, seed=1)
    numbers, positions = numbers.to(dev), positions.to(dev).double()
    charge = torch.zeros(numbers.shape[0], dtype=torch.float64, device=dev)
    eng = eeq._EeqEngine.get(dev)
    q = eng.charges(numbers, positions, charge, 25.0)
    gq = torch.randn_like(q)
    t1 = timed(lambda: eng.charges(numbers, positions, charge, 25.0))
    t2 = timed(lambda: eng.vjp(numbers, positions, 25.0, q, gq))
    line = f"{name:18s} charges {t1 * 1e3:8.1f} us   vjp {t2 * 1e3:8.1f} us"
    if hasattr(eng.lib, "d4b200_eeq_vjp_factor_f64"):
        _, factor = eng.charges(numbers, positions, charge, 25.0, keep_factor=True)
        t3 = timed(lambda: eng.charges(numbers, positions, charge, 25.0, keep_factor=True))
        t4 = timed(lambda: eng.vjp(numbers, positions, 25.0, q, gq, factor))
        line += f"   charges+factor {t3 * 1e3:8.1f} us   vjp from factor {t4 * 1e3:8.1f} us"
    print(line)
