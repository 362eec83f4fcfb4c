"""Small energy / gradient / D4S / properties / large-path calls for compute-sanitizer runs."""
import sys, pathlib
root = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(root)); sys.path.insert(0, str(root / "oracle"))
import torch
import d4_oracle as orc
import tad_dftd4_b200 as d4

dev = torch.device("cuda:0")
param = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)
numbers, positions, q = orc.organic_batch([12, 33, 7, 50, 64, 20], seed=5)
for dtype in (torch.float64, torch.float32):
    n, p, qq = numbers.to(dev), positions.to(dev, dtype), q.to(dev, dtype)
    for model in ("d4", "d4s"):
        pos = p.clone().requires_grad_(True)
        e = d4.dftd4(n, pos, 0.0, param, q=qq, model=model)
        (g,) = torch.autograd.grad(e.sum(), pos)
        e2 = d4.dftd4(n, p, 0.0, param, q=qq, model=model)
        print(dtype, model, float(e.sum()), float(e2.sum()), float(g.abs().max()))
    d4.get_properties(n, p, q=qq)
eh = d4.dftd4_host(numbers, positions, 0.0, param, q=q, chunks=3)
print("host", float(eh.sum()))
eh8, gh8 = d4.dftd4_host(numbers.to(torch.uint8), positions, 0.0, param, q=q, chunks=2, with_gradient=True)
print("host uint8 + gradient", float(eh8.sum()), float(gh8.abs().max()))
# odd padded width: the rows of odd structures are only 8-byte aligned (bulk-async staging windows), views
# that start inside an allocation, many structures per CTA queue, weighted upstream gradient
no, po, qo = orc.organic_batch([12, 33, 7, 21, 5, 30, 33, 2, 19] * 40, seed=6)
for dtype in (torch.float64, torch.float32):
    n, p, qq = no.to(dev), po.to(dev, dtype), qo.to(dev, dtype)
    for sl in (slice(None), slice(1, None), slice(3, -2)):
        pos = p[sl].clone().requires_grad_(True)
        e = d4.dftd4(n[sl], pos, 0.0, param, q=qq[sl])
        w = torch.linspace(0.5, 1.5, e.numel(), device=dev, dtype=dtype).reshape(e.shape)
        (g,) = torch.autograd.grad((e * w).sum(), pos)
        e2 = d4.dftd4(n[sl], p[sl], 0.0, param, q=qq[sl])
        print("odd width", dtype, sl, float(e.sum()), float(e2.sum()), float(g.abs().max()))
nb, pb, qb = orc.organic_batch([150], seed=7)
pos = pb[0].to(dev).requires_grad_(True)
e = d4.dftd4(nb[0].to(dev), pos, 0.0, param, q=qb[0].to(dev))
(g,) = torch.autograd.grad(e.sum(), pos)
print("large", float(e.sum()), float(g.abs().max()))
torch.cuda.synchronize()
