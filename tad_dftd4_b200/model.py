"""
Model descriptors with the constructor signature of the reference's ``D4Model`` /
``D4SModel`` (``/root/reference/src/tad_dftd4/model/base.py:107-151``).

In the reference a model instance owns the ``(N, N, 7, 7)`` reference-C6 tensor of one
set of atoms.  Here that tensor never exists (C6 is a 23-term dot product of per-atom
vectors, see ``tables.py``), so a model is only a bundle of hyper-parameters
``(ga, gc, wf)`` that selects the device tables; it can be passed as ``model=`` to
:func:`tad_dftd4_b200.dftd4` exactly like the reference's instances.
"""

from __future__ import annotations

import torch

from . import defaults

__all__ = ["D4Model", "D4SModel"]


class _Model:
    __slots__ = ("numbers", "ga", "gc", "wf", "ref_charges", "device", "dtype")
    _key = "d4"

    def __init__(self, numbers=None, ga: float = defaults.GA_DEFAULT, gc: float = defaults.GC_DEFAULT,
                 wf=None, ref_charges: str = "eeq", rc6=None, device=None, dtype=None) -> None:  # fmt: skip
        if ref_charges not in ("eeq", "gfn2"):
            raise ValueError(f"Unknown reference charges: {ref_charges}")
        if ref_charges != "eeq":
            raise NotImplementedError("only ref_charges='eeq' is accelerated")
        if rc6 is not None:
            raise NotImplementedError("user-supplied rc6 tensors are outside the accelerated hot path")
        self.numbers = numbers
        self.ga = float(ga)
        self.gc = float(gc)
        self.wf = defaults.WF_DEFAULT if wf is None else float(wf)
        self.ref_charges = ref_charges
        self.device = device
        self.dtype = dtype if dtype is not None else torch.get_default_dtype()

    def __repr__(self) -> str:  # pragma: no cover
        return f"{type(self).__name__}(ga={self.ga}, gc={self.gc}, wf={self.wf}, ref_charges={self.ref_charges})"

    def weight_references(self, cn=None, q=None, **kw):
        raise NotImplementedError(
            "reference weights are evaluated inside the fused kernels; use get_properties() "
            "for coordination numbers, C6 coefficients and polarizabilities"
        )

    get_atomic_c6 = weight_references


class D4Model(_Model):
    """D4 model: one Gaussian weighting factor ``wf`` (default 6)."""


class D4SModel(_Model):
    """D4S model: element-pair specific weighting factors (``data/wfpair.py`` of the reference)."""

    _key = "d4s"
