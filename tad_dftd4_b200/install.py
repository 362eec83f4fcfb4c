"""
``install()``: make ``tad_dftd4`` itself run on the B200 kernels.

Two situations:

* ``tad_dftd4`` is importable (a site that has the reference and its dependencies):
  its entry points ``tad_dftd4.dftd4`` / ``tad_dftd4.disp.dftd4`` /
  ``tad_dftd4.get_properties`` / ``tad_dftd4.disp.get_properties`` are rebound to this
  package's functions and ``tad_dftd4.dispersion.DispD4.calculate`` to the fused path;
  everything else of the reference (parameters, models, I/O) is left alone.
* it is not importable: an alias module named ``tad_dftd4`` that exposes this package's
  mirror of the reference API is registered in ``sys.modules``, so that
  ``import tad_dftd4 as d4`` in unchanged user code resolves to the B200 path.

Unchanged reference scripts (``/root/reference/examples/*.py``) create their tensors on
the CPU.  The kernels have no CPU path, so ``install(device=...)`` sets a *placement
policy*: CPU tensor arguments are copied to that device, the result is copied back to
where ``positions`` lives (both copies are differentiable, so ``autograd.grad`` with
respect to CPU positions works).  Without ``device`` CPU tensors raise as they always do.
"""

from __future__ import annotations

import functools
import importlib
import sys
import types
from typing import Any

import torch

__all__ = ["install", "uninstall", "placed"]

_SAVED: list[tuple[Any, str, Any]] = []
_ALIAS: str | None = None


def _move(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device) if x.device.type == "cpu" else x
    if isinstance(x, dict):  # Param: a plain dict in the reference as well (parameters/base.py:48-85)
        return {k: _move(v, device) for k, v in x.items()}
    return x


def placed(fn, device):
    """Wrap ``fn`` with the placement policy described in the module docstring."""
    if device is None:
        return fn
    device = torch.device(device)

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        home = None
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_floating_point() and a.dim() >= 2:
                home = a.device  # positions
                break
        args = tuple(_move(a, device) for a in args)
        kwargs = {k: _move(v, device) for k, v in kwargs.items()}
        out = fn(*args, **kwargs)
        if home is None or home == device:
            return out
        if isinstance(out, tuple):
            return tuple(o.to(home) if isinstance(o, torch.Tensor) else o for o in out)
        return out.to(home)

    return wrapper


def _set(obj, name, value):
    _SAVED.append((obj, name, getattr(obj, name, None)))
    setattr(obj, name, value)


def install(device: str | torch.device | int | None = None, name: str = "tad_dftd4") -> str:
    """Route ``tad_dftd4``'s entry points to the B200 kernels.  Returns ``"rebound"`` when the
    reference package was found and patched, ``"alias"`` when an alias module was registered."""
    global _ALIAS
    import tad_dftd4_b200 as pkg
    from tad_dftd4_b200 import dispersion as disp_mod

    if isinstance(device, int):
        device = torch.device("cuda", device)
    uninstall()
    dftd4 = placed(pkg.dftd4, device)
    get_properties = placed(pkg.get_properties, device)
    try:
        ref = importlib.import_module(name)
        if getattr(ref, "__d4b200_alias__", False):
            raise ImportError
    except ImportError:
        ref = None
    if ref is not None:
        _set(ref, "dftd4", dftd4)
        _set(ref, "get_properties", get_properties)
        sub = getattr(ref, "disp", None)
        if isinstance(sub, types.ModuleType):
            _set(sub, "dftd4", dftd4)
            _set(sub, "get_properties", get_properties)
        rdisp = getattr(ref, "dispersion", None)
        for cls_name in ("DispD4", "DispD4Exact"):  # DispD4Exact lives in dispersion.d4 of the reference
            owner = rdisp if hasattr(rdisp, cls_name) else getattr(rdisp, "d4", None)
            if owner is None or not hasattr(owner, cls_name):
                continue
            mine_cls = getattr(disp_mod, cls_name)

            def calculate(self, numbers, positions, charge, param, _cls=mine_cls, **kw):
                # the reference instance keeps its model as a key (+ kwargs) or as an instance (base.py:152-178)
                inst = getattr(self, "_model_instance", None)
                if inst is not None:
                    mine = _cls(model=inst)
                else:
                    mine = _cls(model=getattr(self, "_model_key", "d4"),
                                model_kwargs=getattr(self, "_model_kwargs", None) or None)
                return placed(mine.calculate, device)(numbers, positions, charge, param, **kw)

            _set(getattr(owner, cls_name), "calculate", calculate)
        return "rebound"
    alias = types.ModuleType(name)
    alias.__d4b200_alias__ = True
    alias.__doc__ = f"alias of tad_dftd4_b200 registered by tad_dftd4_b200.install.install() as {name!r}"
    for key in pkg.__all__:
        setattr(alias, key, getattr(pkg, key))
    alias.dftd4 = dftd4
    alias.get_properties = get_properties
    if device is not None:
        # class-based interface under the same placement policy
        dsp = types.ModuleType(name + ".dispersion")
        for key in disp_mod.__all__:
            setattr(dsp, key, getattr(disp_mod, key))

        class DispD4(disp_mod.DispD4):
            def calculate(self, *a, **kw):  # noqa: D102
                return placed(super().calculate, device)(*a, **kw)

        class DispD4Exact(disp_mod.DispD4Exact):
            def calculate(self, *a, **kw):  # noqa: D102
                return placed(super().calculate, device)(*a, **kw)

        dsp.DispD4, dsp.DispD4Exact = DispD4, DispD4Exact
        alias.dispersion = dsp
        sys.modules[name + ".dispersion"] = dsp
    dmod = types.ModuleType(name + ".disp")
    dmod.dftd4, dmod.get_properties = dftd4, get_properties
    alias.disp = dmod
    sys.modules[name + ".disp"] = dmod
    sys.modules[name] = alias
    _ALIAS = name
    return "alias"


def uninstall() -> None:
    """Undo :func:`install`."""
    global _ALIAS
    while _SAVED:
        obj, name, old = _SAVED.pop()
        if old is None:
            try:
                delattr(obj, name)
            except AttributeError:
                pass
        else:
            setattr(obj, name, old)
    if _ALIAS is not None:
        for key in (_ALIAS, _ALIAS + ".disp", _ALIAS + ".dispersion"):
            mod = sys.modules.get(key)
            if mod is not None and (getattr(mod, "__d4b200_alias__", False) or key != _ALIAS):
                sys.modules.pop(key, None)
        _ALIAS = None
