"""``any_to_tensor`` (restated)."""
from __future__ import annotations

import torch

__all__ = ["any_to_tensor", "symbol_to_number", "reshape_fortran"]


def any_to_tensor(x, device=None, dtype=None):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype)
    if isinstance(x, (bool, int, float, list, tuple)):
        return torch.tensor(x, device=device, dtype=dtype)
    raise TypeError(f"Cannot convert {type(x)} to tensor.")


_PSE = (
    "X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co "
    "Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te "
    "I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir "
    "Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No "
    "Lr Rf Db Sg Bh Hs Mt Ds Rg Cn Nh Fl Mc Lv Ts Og"
).split()


def symbol_to_number(symbols):
    return torch.tensor([_PSE.index(s.capitalize()) for s in symbols])


def reshape_fortran(x, shape):
    if len(x.shape) > 0:
        x = x.permute(*reversed(range(len(x.shape))))
    return x.reshape(*reversed(shape)).permute(*reversed(range(len(shape))))
