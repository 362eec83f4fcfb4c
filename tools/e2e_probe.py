"""e2e breakdown of the host-buffer energy call (development probe; GPU box):
chunk-count sweep, pure H2D/D2H copy time of the same bytes, resident kernel time."""
import statistics, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import torch
import bench
import tad_dftd4_b200 as d4

wl = bench.WORKLOADS["c2"]
numbers_h, positions_h, q_h = bench.make_batch(wl, 0)
numbers_h, positions_h, q_h = numbers_h.pin_memory(), positions_h.pin_memory(), q_h.pin_memory()
dev = torch.device("cuda:0")
numbers, positions, q = numbers_h.to(dev), positions_h.to(dev), q_h.to(dev)
out = torch.empty(numbers_h.shape, dtype=torch.float64).pin_memory()
d4.set_checks(False)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

def wall(fn, reps=40):
    ts = []
    for _ in range(5): fn()
    torch.cuda.synchronize()
    for r in range(reps):
        flush.fill_(float(r)); torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts), min(ts)

def copies():
    a = numbers_h.to(dev, non_blocking=True); b = positions_h.to(dev, non_blocking=True); c = q_h.to(dev, non_blocking=True)
print("H2D only (9.8 MB): median %.3f min %.3f ms" % wall(copies))
e_dev = torch.empty(numbers.shape, dtype=torch.float64, device=dev)
print("D2H only (2 MB): median %.3f min %.3f ms" % wall(lambda: out.copy_(e_dev, non_blocking=True)))
print("resident kernels: median %.3f min %.3f ms" % wall(lambda: d4.dftd4(numbers, positions, 0.0, bench.PBE0, q=q)))
for ch in (1, 2, 4, 6, 8):
    m, lo = wall(lambda: d4.dftd4_host(numbers_h, positions_h, 0.0, bench.PBE0, q=q_h, device=dev, out=out, chunks=ch))
    print("dftd4_host chunks=%2d: median %.3f min %.3f ms" % (ch, m, lo))
