"""Per-phase cycle breakdown of the small-family kernels (development probe; GPU box)."""
import ctypes as C, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import torch
import bench
import tad_dftd4_b200 as d4
from tad_dftd4_b200 import _lib
from tad_dftd4_b200.disp import _Engine

wlname = sys.argv[1] if len(sys.argv) > 1 else "c2"
wl = bench.WORKLOADS[wlname]
numbers, positions, q = bench.make_batch(wl, 0)
dev = torch.device("cuda:0")
numbers, positions, q = numbers.to(dev), positions.to(dev), q.to(dev)
d4.set_checks(False)
def step():
    if wl["grad"]:
        pos = positions.detach().requires_grad_(True)
        e = d4.dftd4(numbers, pos, 0.0, bench.PBE0, q=q)
        torch.autograd.grad(e.sum(), pos)
    else:
        d4.dftd4(numbers, positions, 0.0, bench.PBE0, q=q)
for _ in range(3): step()
torch.cuda.synchronize()
eng = _Engine.get(dev, 3.0, 2.0); lib = _lib.load()
lib.d4b200_phase_profile(eng.handle, 1, None)
reps = 5
for _ in range(reps): step()
out = (C.c_ulonglong * (5 * 16))()
lib.d4b200_phase_profile(eng.handle, 0, out)
names = ["load", "cn pairs", "cn rows", "weights", "A vec", "e2 pairs", "A0 vec", "stash", "triples", "final",
         "g:coef", "g:B", "g:proj", "g:cnchain", "g:forces", "-"]
for c in range(5):
    row = [out[c * 16 + k] for k in range(16)]
    tot = sum(row)
    if tot == 0: continue
    print(f"class {c}: total {tot/reps/1e6:.2f} Mcycles/step (sum over CTAs)")
    print("   " + "  ".join(f"{names[k]} {100*row[k]/tot:.1f}%" for k in range(16) if row[k]))
