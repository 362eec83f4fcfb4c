"""
Real-space cutoffs: same constructor, attributes and defaults as the
reference's ``Cutoff`` (``/root/reference/src/tad_dftd4/cutoff.py:36-90``).

``cn`` is carried for API compatibility only: the reference never forwards it
to the coordination number (``dispersion/base.py:390``), which therefore always
uses 30 Bohr; this package reproduces that.
"""

from __future__ import annotations

import torch

from . import defaults

__all__ = ["Cutoff"]


def _t(x, device, dtype):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype)
    return torch.tensor(x, device=device, dtype=dtype)


def _f(x):
    """Python float of a cutoff given as a number (tensors are converted lazily,
    a device tensor costs one synchronisation)."""
    return None if isinstance(x, torch.Tensor) else float(x)


class Cutoff:
    """Collection of real-space cutoffs (Bohr)."""

    __slots__ = ("disp2", "disp3", "cn", "cn_eeq", "_device", "_dtype", "_floats")

    def __init__(
        self,
        disp2=defaults.D4_DISP2_CUTOFF,
        disp3=defaults.D4_DISP3_CUTOFF,
        cn=defaults.D4_CN_CUTOFF,
        cn_eeq=defaults.D4_CN_EEQ_CUTOFF,
        device: torch.device | None = None,
        dtype: torch.dtype | None = None,
    ) -> None:
        self._device = device if device is not None else torch.device("cpu")
        self._dtype = dtype if dtype is not None else torch.get_default_dtype()
        self.disp2 = _t(disp2, device, dtype)
        self.disp3 = _t(disp3, device, dtype)
        self.cn = _t(cn, device, dtype)
        self.cn_eeq = _t(cn_eeq, device, dtype)
        self._floats = {"disp2": _f(disp2), "disp3": _f(disp3), "cn": _f(cn), "cn_eeq": _f(cn_eeq)}

    def as_float(self, name: str) -> float:
        """Host value of a cutoff without touching the device when it was given
        as a Python number."""
        v = self._floats.get(name)
        if v is None:
            v = self._floats[name] = float(getattr(self, name))
        return v

    @property
    def device(self):
        return self._device

    @property
    def dtype(self):
        return self._dtype

    @property
    def dd(self):
        return {"device": self._device, "dtype": self._dtype}

    def type(self, dtype: torch.dtype) -> "Cutoff":
        if dtype not in (torch.float16, torch.float32, torch.float64):
            raise ValueError(f"Only float types are allowed (got {dtype}).")
        return Cutoff(self.disp2, self.disp3, self.cn, self.cn_eeq, device=self._device, dtype=dtype)

    def to(self, device: torch.device) -> "Cutoff":
        return Cutoff(self.disp2, self.disp3, self.cn, self.cn_eeq, device=device, dtype=self._dtype)
