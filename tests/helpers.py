"""Shared helpers for the test-suite: golden fixture loading."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"
GOLDEN_CASES = sorted(p.stem for p in GOLDEN.glob("*.npz"))


def load_golden(name: str) -> dict:
    raw = np.load(GOLDEN / f"{name}.npz")
    out = {k: raw[k] for k in raw.files}
    out["param"] = {str(k): float(v) for k, v in zip(out.pop("param_keys"), out.pop("param_vals"))}
    out["cutoff"] = {str(k): float(v) for k, v in zip(out.pop("cutoff_keys"), out.pop("cutoff_vals"))}
    return out


def as_torch(case: dict, device="cpu", dtype=torch.float64):
    numbers = torch.from_numpy(case["numbers"]).to(device)
    positions = torch.from_numpy(case["positions"]).to(device=device, dtype=dtype)
    q = torch.from_numpy(case["q"]).to(device=device, dtype=dtype)
    return numbers, positions, q
