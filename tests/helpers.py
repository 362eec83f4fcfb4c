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


PARAM_KEYS = ("s6", "s8", "s9", "s10", "a1", "a2", "alp")
PARAM_GOLDEN_CASES = sorted(p.stem for p in (GOLDEN / "param").glob("*.npz"))


def load_param_golden(name: str) -> dict:
    """Parameter-gradient fixture of ``oracle/make_golden_param.py`` (unmodified reference): the
    inputs are those of ``tests/golden/<name>.npz``; adds ``pvalues`` (the seven damping
    parameters), ``g`` (upstream weights) and ``grad_param_d4`` / ``grad_param_d4s``."""
    case = load_golden(name)
    raw = np.load(GOLDEN / "param" / f"{name}.npz")
    case["pvalues"] = {str(k): float(v) for k, v in zip(raw["param_keys"], raw["param_vals"])}
    case["g"] = raw["g"]
    case["grad_param_d4"] = raw["grad_param_d4"]
    case["grad_param_d4s"] = raw["grad_param_d4s"]
    return case
