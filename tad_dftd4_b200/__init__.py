"""
tad_dftd4_b200 -- B200-native (sm_100a) implementation of the DFT-D4 hot path
behind the functional API of ``tad_dftd4``::

    import tad_dftd4_b200 as d4
    energy = d4.dftd4(numbers, positions, charge, param, q=q)   # (..., nat)

CUDA kernels: ``csrc/`` (C ABI in ``include/d4b200.h``).  No CPU fallback.
"""

from . import batch, cutoff, damping, defaults, dispersion, eeq, large, model, ncoord, parallel
from .batch import pack
from .cutoff import Cutoff
from .damping import MZeroDamping, OptimisedPowerDamping, Param, RationalDamping, ZeroDamping, get_params
from .eeq import get_eeq_charges
from .disp import dftd4, dftd4_host, get_properties, last_launch_count, set_checks, set_fused_forward
from .install import install, uninstall
from .model import D4Model, D4SModel

__version__ = "0.1.0"

__all__ = [
    "__version__", "batch", "cutoff", "Cutoff", "damping", "defaults", "dftd4", "dftd4_host", "get_params",
    "get_properties", "pack", "Param", "RationalDamping", "set_checks", "set_fused_forward", "last_launch_count",
    "ZeroDamping", "MZeroDamping", "OptimisedPowerDamping", "dispersion", "eeq", "get_eeq_charges", "large", "model", "ncoord", "parallel", "D4Model", "D4SModel",
    "install", "uninstall",
]  # fmt: skip
