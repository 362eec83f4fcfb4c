"""Element tables as callables ``TABLE(device=None, dtype=None) -> Tensor``."""
from __future__ import annotations

import importlib.util
from pathlib import Path

import torch

from . import zeff  # noqa: F401
from ._tables import COV_D3, GAM, PAULING, VDW_PAIRWISE, ZEFF

__all__ = ["COV_D3", "GAM", "PAULING", "VDW_PAIRWISE", "ZEFF"]
