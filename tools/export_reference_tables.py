#!/usr/bin/env python
"""
Export the DFT-D4 parameter DATA (no code) from the read-only reference tree
into the two data files the package ships:

* ``tad_dftd4_b200/data/d4_reference.npz`` -- literal parameter arrays of
  ``/root/reference/src/tad_dftd4/reference/d4/params.py`` (refcovcn, refalpha,
  refascale, refscount, refsys, refc, secscale, secalpha),
  ``reference/d4/charge_eeq.py`` (clsq, clsh), ``reference/d4/charge_gfn2.py`` (refq, refh),
  ``data/r4r2.py`` (already
  transformed, ``r4r2.py:83-88``) and ``data/wfpair.py`` (119x119, D4S).
* ``tad_dftd4_b200/data/d4_damping.json`` -- the rational-damping parameter
  blocks of ``damping/parameters/d4.toml``.

The data modules only need ``torch``; they are imported by file path.  Run in
the build container only (``/root/reference`` does not exist on the GPU box).
"""

from __future__ import annotations

import importlib.util
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/src/tad_dftd4")
OUT = Path(__file__).resolve().parent.parent / "tad_dftd4_b200" / "data"


def _load(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main() -> None:
    # wfpair.py imports `Tensor` from tad_mctc.typing only for annotations
    if "tad_mctc" not in sys.modules:
        pkg = types.ModuleType("tad_mctc")
        typ = types.ModuleType("tad_mctc.typing")
        typ.Tensor = torch.Tensor
        pkg.typing = typ
        sys.modules["tad_mctc"] = pkg
        sys.modules["tad_mctc.typing"] = typ

    params = _load("_ref_params", REF / "reference/d4/params.py")
    ceeq = _load("_ref_charge_eeq", REF / "reference/d4/charge_eeq.py")
    cgfn = _load("_ref_charge_gfn2", REF / "reference/d4/charge_gfn2.py")
    r4r2 = _load("_ref_r4r2", REF / "data/r4r2.py")
    wfp = _load("_ref_wfpair", REF / "data/wfpair.py")

    arrays = {
        "refcovcn": params.refcovcn.numpy().astype(np.float64),
        "refalpha": params.refalpha.numpy().astype(np.float64),
        "refascale": params.refascale.numpy().astype(np.float64),
        "refscount": params.refscount.numpy().astype(np.float64),
        "refsys": params.refsys.numpy().astype(np.int32),
        "refc": params.refc.numpy().astype(np.int32),
        "secscale": params.secscale.numpy().astype(np.float64),
        "secalpha": params.secalpha.numpy().astype(np.float64),
        "clsq": ceeq.clsq.numpy().astype(np.float64),
        "clsh": ceeq.clsh.numpy().astype(np.float64),
        # ref_charges="gfn2" (model/base.py:393-397, model/d4.py:145-147)
        "gfn2_refq": cgfn.refq.numpy().astype(np.float64),
        "gfn2_refh": cgfn.refh.numpy().astype(np.float64),
        # r4r2.py:83-88 evaluated in float64
        "r4r2": r4r2.R4R2(dtype=torch.float64).numpy(),
        "wfpair": wfp.WFPAIR(dtype=torch.float64).numpy(),
    }
    for k, v in arrays.items():
        print(f"{k:10s} {str(v.shape):16s} {v.dtype}")
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / "d4_reference.npz", **arrays)

    import tomli

    with open(REF / "damping/parameters/d4.toml", "rb") as fp:
        table = tomli.load(fp)
    with open(OUT / "d4_damping.json", "w", encoding="utf8") as fp:
        json.dump(table, fp, indent=1, sort_keys=True)
    print("functionals:", len(table["parameter"]))


if __name__ == "__main__":
    main()
