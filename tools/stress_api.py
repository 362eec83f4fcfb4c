"""Randomised sequence of public-API calls in ONE process, each checked against the float64 oracle.

State that survives a call (engines, workspaces per stream, the plan cache of the tiled family, cached EEQ factors,
exit hooks) is what this is after: sizes, models, dtypes, gradients, charge sources and table variants change from call
to call, freed tensors hand their addresses to the next case.

    python tools/stress_api.py [ncases] [seed]
"""
import pathlib
import random
import sys

import torch

root = pathlib.Path(__file__).resolve().parents[1]
sys.path[:0] = [str(root), str(root / "oracle")]
import d4_oracle as orc  # noqa: E402
import eeq_oracle  # noqa: E402

import tad_dftd4_b200 as d4  # noqa: E402
from tad_dftd4_b200.dispersion import DispD4Exact  # noqa: E402

dev = torch.device("cuda:0")
ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
PAR = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)
worst = {"e64": 0.0, "g64": 0.0, "e32": 0.0}
fails = 0
for case in range(ncases):
    kind = rng.choice(["batch", "batch", "batch", "large", "exact", "gfn2", "eeq"])
    model = rng.choice(["d4", "d4s"]) if kind in ("batch", "gfn2", "eeq") else "d4"
    f32 = kind == "batch" and rng.random() < 0.3
    grad = rng.random() < 0.6
    if kind == "large":
        sizes = [rng.choice([150, 150, 170, rng.randint(130, 260)])]  # repeated sizes: freed addresses come back
    elif kind == "exact":
        sizes = [rng.randint(2, 60) for _ in range(rng.randint(1, 4))]
    else:
        sizes = [rng.randint(1, 100 if grad or model == "d4s" else 120) for _ in range(rng.randint(1, 12))]
    numbers, positions, q = orc.organic_batch(sizes, seed=rng.randint(0, 10**6))
    if kind == "large":
        numbers, positions, q = numbers[0], positions[0], q[0]
    okw = {}
    if kind == "gfn2":
        okw["ref_charges"] = "gfn2"
    if kind == "exact":
        okw["c9"] = "exact"
    q_ref = q
    if kind == "eeq":
        chg = torch.zeros(numbers.shape[:-1], dtype=torch.float64)
        pr = positions.clone().requires_grad_(grad)
        q_ref = eeq_oracle.get_eeq_charges(numbers, pr, chg)
        e_ref = orc.dftd4(numbers, pr, PAR, q_ref, model=model)
        g_ref = torch.autograd.grad(e_ref.sum(), pr)[0] if grad else None
        e_ref = e_ref.detach()
    elif grad:
        e_ref, g_ref = orc.energy_and_gradient(numbers, positions, PAR, q, model=model, **okw)
    else:
        e_ref, g_ref = orc.dftd4(numbers, positions, PAR, q, model=model, **okw), None
    dt = torch.float32 if f32 else torch.float64
    n_d = numbers.to(dev)
    pos = positions.to(dev, dt).requires_grad_(grad)
    q_d = None if kind == "eeq" else q.to(dev, dt)
    if kind == "exact":
        e = DispD4Exact().calculate(n_d, pos, 0.0, PAR, q=q_d)
    elif kind == "gfn2":
        cls = d4.D4Model if model == "d4" else d4.D4SModel
        e = d4.dftd4(n_d, pos, 0.0, PAR, q=q_d, model=cls(n_d, ref_charges="gfn2"))
    else:
        e = d4.dftd4(n_d, pos, 0.0, PAR, q=q_d, model=model)
    de = ((e.detach().cpu().double() - e_ref).abs().max() / e_ref.abs().max().clamp(min=1e-300)).item()
    dg = 0.0
    if grad:
        (g,) = torch.autograd.grad(e.sum(), pos)
        dg = (g.cpu().double() - g_ref).abs().max().item()
    tol_e, tol_g = (1e-5, 1e-5) if f32 else (1e-10, 1e-9)
    ok = de < tol_e and dg < tol_g
    fails += not ok
    if f32:
        worst["e32"] = max(worst["e32"], de)
    else:
        worst["e64"], worst["g64"] = max(worst["e64"], de), max(worst["g64"], dg)
    print(f"{case:3d} {kind:5s} {model:3s} {'f32' if f32 else 'f64'} grad={int(grad)} sizes={sizes}  dE {de:.1e} dG {dg:.1e} "
          f"{'ok' if ok else 'FAIL'}")
    del e, pos, n_d, q_d
print(f"{ncases} cases, {fails} failures; worst FP64 dE {worst['e64']:.1e} dG {worst['g64']:.1e}, worst FP32 dE {worst['e32']:.1e}")
sys.exit(1 if fails else 0)
