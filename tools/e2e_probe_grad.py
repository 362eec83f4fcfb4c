"""chunk-count sweep of the host-buffer energy+forces call on the C3 workload (development probe; GPU box)."""
import statistics, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import torch
import bench
import tad_dftd4_b200 as d4

wl = bench.WORKLOADS["c3"]
numbers_h, positions_h, q_h = bench.make_batch(wl, 0)
numbers_h, positions_h, q_h = numbers_h.pin_memory(), positions_h.pin_memory(), q_h.pin_memory()
dev = torch.device("cuda:0")
numbers, positions, q = numbers_h.to(dev), positions_h.to(dev), q_h.to(dev)
out = torch.empty(numbers_h.shape, dtype=torch.float64).pin_memory()
outg = torch.empty(positions_h.shape, dtype=torch.float64).pin_memory()
d4.set_checks(False)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

def wall(fn, reps=25):
    ts = []
    for _ in range(4): fn()
    torch.cuda.synchronize()
    for r in range(reps):
        flush.fill_(float(r)); torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts), min(ts)

def resident():
    pos = positions.detach().requires_grad_(True)
    e = d4.dftd4(numbers, pos, 0.0, bench.PBE0, q=q)
    torch.autograd.grad(e.sum(), pos)
print("resident kernels: median %.3f min %.3f ms" % wall(resident))
for ch in (1, 2, 3, 4, 6, 8, 12):
    m, lo = wall(lambda: d4.dftd4_host(numbers_h, positions_h, 0.0, bench.PBE0, q=q_h, device=dev, out=out,
                                       with_gradient=True, out_gradient=outg, chunks=ch))
    print("dftd4_host forces chunks=%2d: median %.3f min %.3f ms" % (ch, m, lo))
