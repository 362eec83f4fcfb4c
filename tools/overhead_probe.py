"""Where does the per-call time go?  (development probe, run on the GPU box)"""
import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, torch
import bench
import tad_dftd4_b200 as d4

wl = bench.WORKLOADS["c2"]
numbers, positions, q = bench.make_batch(wl, 0)
dev = torch.device("cuda:0")
numbers, positions, q = numbers.to(dev), positions.to(dev), q.to(dev)
d4.set_checks(False)
for _ in range(5):
    d4.dftd4(numbers, positions, 0.0, bench.PBE0, q=q)
torch.cuda.synchronize()
# back-to-back calls, no flush: device time per call
for reps in (1, 20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        d4.dftd4(numbers, positions, 0.0, bench.PBE0, q=q)
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"reps={reps}: device {e0.elapsed_time(e1)/reps*1e3:.1f} us/call, host enqueue {(t1-t0)/reps*1e6:.1f} us/call")
