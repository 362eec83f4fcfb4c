"""Typing helpers + ``TensorLike`` base (restated; see package docstring)."""
from __future__ import annotations

from typing import (  # noqa: F401
    Any,
    Callable,
    Literal,
    NoReturn,
    Protocol,
    TypedDict,
    overload,
)

import torch
from torch import Tensor  # noqa: F401
from typing_extensions import TypeAlias  # noqa: F401


class DD(TypedDict):
    device: "torch.device | None"
    dtype: "torch.dtype | None"


CountingFunction = Callable[..., Tensor]
CNFunc = Callable[..., Tensor]


class Molecule(TypedDict):
    numbers: Tensor
    positions: Tensor


class TensorLike:
    """Base class carrying ``device``/``dtype`` with ``.to`` / ``.type``."""

    __slots__ = ["__device", "__dtype"]
    allowed_dtypes = (torch.float16, torch.float32, torch.float64)

    def __init__(self, device=None, dtype=None):
        self.__device = (
            device if device is not None else torch.tensor(0.0).device
        )
        self.__dtype = dtype if dtype is not None else torch.get_default_dtype()

    @property
    def device(self):
        return self.__device

    @device.setter
    def device(self, *_: Any) -> NoReturn:
        raise AttributeError("Move object to device using the `.to` method")

    @property
    def dtype(self):
        return self.__dtype

    @dtype.setter
    def dtype(self, *_: Any) -> NoReturn:
        raise AttributeError("Change object dtype using the `.type` method")

    @property
    def dd(self) -> DD:
        return {"device": self.device, "dtype": self.dtype}

    def _all_slots(self):
        names = []
        for cls in type(self).__mro__:
            for s in getattr(cls, "__slots__", ()):
                if not s.startswith("__"):
                    names.append(s)
        return names

    def _convert(self, fn, **new):
        args = {}
        for s in self._all_slots():
            if not hasattr(self, s):
                continue
            attr = getattr(self, s)
            if isinstance(attr, Tensor) or issubclass(type(attr), TensorLike):
                attr = fn(attr)
            args[s] = attr
        dd = {"device": self.device, "dtype": self.dtype}
        dd.update(new)
        obj = self.__class__.__new__(self.__class__)
        TensorLike.__init__(obj, **dd)
        for k, v in args.items():
            object.__setattr__(obj, k, v)
        return obj

    def type(self, dtype):
        if self.dtype == dtype:
            return self
        if dtype not in self.allowed_dtypes:
            raise ValueError(f"Only float types are allowed (got {dtype}).")

        def _cast(x):
            if isinstance(x, Tensor) and not x.is_floating_point():
                return x
            return x.type(dtype)

        return self._convert(_cast, dtype=dtype)

    def to(self, device):
        if self.device == device:
            return self
        return self._convert(lambda x: x.to(device), device=device)
