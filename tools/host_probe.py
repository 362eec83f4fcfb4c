"""Timing probe for the host-buffer entry point: chunk-count sweep + raw copy bandwidth."""
import sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
import bench
import tad_dftd4_b200 as d4

wl = bench.WORKLOADS["c2"]
numbers, positions, q = bench.make_batch(wl, 0)
numbers, positions, q = numbers.pin_memory(), positions.pin_memory(), q.pin_memory()
out = torch.empty(numbers.shape, dtype=positions.dtype).pin_memory()
dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

def wall(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for i in range(n):
        flush.fill_(float(i)); torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3

pd = positions.to(dev)
print("h2d positions (5.9MB) ms", wall(lambda: pd.copy_(positions, non_blocking=True)))
nd, qd = numbers.to(dev), q.to(dev)
print("h2d all three ms", wall(lambda: (nd.copy_(numbers, non_blocking=True), pd.copy_(positions, non_blocking=True), qd.copy_(q, non_blocking=True))))
print("resident ms", wall(lambda: d4.dftd4(nd, pd, 0.0, bench.PBE0, q=qd)))
for ch in (1, 2, 3, 4, 6, 8, 12, 16):
    print("chunks", ch, "ms (median, min)", wall(lambda: d4.dftd4_host(numbers, positions, 0.0, bench.PBE0, q=q, device=dev, out=out, chunks=ch)))
