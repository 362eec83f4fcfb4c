#!/usr/bin/env python
"""
TEST INFRASTRUCTURE ONLY -- generates ``tests/golden/*.npz``.

Runs the UNMODIFIED reference (``/root/reference/src/tad_dftd4``, imported on
top of ``oracle/mctc_shim`` because its dependency ``tad-mctc`` is not
installable here) in float64 on a fixed set of cases, stores inputs and
outputs as small fixtures and, in the same run, checks that the self-contained
restatement ``oracle/d4_oracle.py`` reproduces every stored number (energies
to <= 1e-12 relative, gradients to <= 1e-13 absolute).

Run in the build container only:  ``python oracle/make_golden.py``
(``/root/reference`` does not exist on the GPU box; the fixtures travel).
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "mctc_shim"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, str(HERE))

import tad_dftd4 as ref  # noqa: E402  (the real reference)
from tad_mctc.ncoord import cn_d4 as ref_cn_d4  # noqa: E402  (shim)

import d4_oracle as orc  # noqa: E402

OUT = HERE.parent / "tests" / "golden"
F64 = torch.float64

TPSSH = dict(s6=1.0, s8=1.85897750, s9=1.0, a1=0.44286966, a2=4.60230534)  # examples/single.py:31-37
PBE0 = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)  # d4.toml:269
TPSS0 = dict(s6=1.0, s8=1.62438102, s9=1.0, a1=0.40329022, a2=4.80537871)  # test_grad/test_pos.py:51-57

SINGLE_Z = [6, 6, 6, 6, 7, 6, 16, 1, 1, 1, 1, 1]  # examples/single.py:7-9
SINGLE_XYZ = [
    [-2.56745685564671, -0.02509985979910, 0.0],
    [-1.39177582455797, +2.27696188880014, 0.0],
    [+1.27784995624894, +2.45107479759386, 0.0],
    [+2.62801937615793, +0.25927727028120, 0.0],
    [+1.41097033661123, -1.99890996077412, 0.0],
    [-1.17186102298849, -2.34220576284180, 0.0],
    [-2.39505990368378, -5.22635838332362, 0.0],
    [+2.41961980455457, -3.62158019253045, 0.0],
    [-2.51744374846065, +3.98181713686746, 0.0],
    [+2.24269048384775, +4.24389473203647, 0.0],
    [+4.66488984573956, +0.17907568006409, 0.0],
    [-4.60044244782237, -0.17794734637413, 0.0],
]
NAN17_Z = [6, 6, 6, 6, 6, 6, 6, 6, 1, 1, 1, 1, 1, 7, 8, 8, 8]  # test/test_grad/test_nan.py:37-58
NAN17_XYZ = [
    [-1.0981, +0.1496, +0.1346], [-0.4155, +1.2768, +0.3967], [+0.9426, +0.7848, +0.1307],
    [+2.1708, +1.3814, -0.0347], [+3.3234, +0.5924, -0.1535], [+3.1564, -0.8110, -0.0285],
    [+1.8929, -1.4673, +0.0373], [+0.8498, -0.5613, +0.0109], [-0.7751, +2.2970, +0.5540],
    [+2.3079, +2.4725, -0.1905], [+4.3031, +0.9815, -0.4599], [+4.0011, -1.4666, -0.0514],
    [+1.8340, -2.5476, -0.1587], [-2.5629, -0.0306, -0.1458], [-3.0792, +1.0280, -0.3225],
    [-3.0526, -1.1594, +0.1038], [-0.4839, -0.9612, -0.0048],
]  # fmt: skip


def zero_sum_charges(numbers: np.ndarray, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    q = 0.1 * rng.normal(size=numbers.shape) * (numbers != 0)
    for b in range(q.shape[0]) if q.ndim == 2 else [None]:
        row = q if b is None else q[b]
        real = (numbers if b is None else numbers[b]) != 0
        if real.any():
            row[real] -= row[real].mean()
    return q


def cases():
    """name -> (numbers, positions, q, param, kwargs)"""
    out = {}
    a = 1.61768389755830
    out["sih4_tpssh"] = (
        np.array([14, 1, 1, 1, 1]),
        np.array([[0, 0, 0], [a, a, -a], [-a, -a, -a], [a, -a, a], [-a, a, a]], dtype=float),
        # test/test_d4/samples.py:88-99
        np.array([-8.412842390895063e-02, 2.103210597723753e-02, 2.103210597723774e-02,
                  2.103210597723764e-02, 2.103210597723773e-02]),
        TPSSH, {},
    )  # fmt: skip
    z = 1.50796743897235
    out["lih_tpssh"] = (
        np.array([3, 1]),
        np.array([[0, 0, -z], [0, 0, z]], dtype=float),
        np.array([3.708714958301688e-01, -3.708714958301688e-01]),  # samples.py:53-56
        TPSSH, {},
    )
    zs = np.array(SINGLE_Z)
    out["single_pbe0"] = (zs, np.array(SINGLE_XYZ), zero_sum_charges(zs, 11), PBE0, {})
    out["single_tpssh_s10"] = (
        zs, np.array(SINGLE_XYZ), zero_sum_charges(zs, 11), dict(TPSSH, s10=0.7, alp=14.0), {},
    )  # fmt: skip
    zn = np.array(NAN17_Z)
    out["nan17"] = (
        zn, np.array(NAN17_XYZ), zero_sum_charges(zn, 12),
        dict(s6=1.0, s8=0.78981345, s9=1.0, a1=0.49484001, a2=5.73083694), {},
    )  # fmt: skip

    for nat, seed in [(2, 101), (3, 102), (5, 103), (20, 104), (33, 105), (64, 106), (65, 107), (100, 108), (128, 109)]:
        zz, xyz, q = orc.organic_blob(nat, np.random.default_rng(seed))
        out[f"organic_{nat}"] = (zz, xyz, q, PBE0, {})

    # ragged padded batch incl. a single atom and an all-padding structure
    n, p, q = orc.organic_batch([20, 60, 33, 1, 47, 2], seed=7)
    n, p, q = n.numpy(), p.numpy(), q.numpy()
    n = np.concatenate([n, np.zeros((1, n.shape[1]), dtype=n.dtype)])
    p = np.concatenate([p, np.zeros((1,) + p.shape[1:])])
    q = np.concatenate([q, np.zeros((1, q.shape[1]))])
    out["ragged_batch"] = (n, p, q, TPSS0, {})

    # padding in the middle of the atom axis
    zz, xyz, q = orc.organic_blob(24, np.random.default_rng(21))
    zz2 = np.zeros(30, dtype=zz.dtype)
    xyz2 = np.zeros((30, 3))
    q2 = np.zeros(30)
    keep = np.array([0, 1, 2, 4, 5, 6, 7, 9, 10, 11, 12, 13, 15, 16, 18, 19, 20, 21, 22, 24, 25, 26, 28, 29])
    zz2[keep], xyz2[keep], q2[keep] = zz, xyz, q
    out["holes"] = (zz2, xyz2, q2, PBE0, {})

    # all elements of the reference tables on a jittered lattice
    rng = np.random.default_rng(31)
    zz = np.arange(1, 104)
    grid = np.stack(np.meshgrid(*[np.arange(5)] * 3, indexing="ij"), -1).reshape(-1, 3)[:103]
    xyz = grid * 7.5 + rng.normal(scale=0.4, size=(103, 3))
    out["all_elements"] = (zz, xyz, zero_sum_charges(zz, 32) * 3.0, TPSSH, {})

    # 150 Bohr zig-zag chain: both hard cutoffs and the ATM two-distance mask bite
    rng = np.random.default_rng(41)
    nat = 48
    zz = np.where(np.arange(nat) % 3 == 0, 6, np.where(np.arange(nat) % 3 == 1, 8, 1))
    xyz = np.stack([np.arange(nat) * 3.2, (np.arange(nat) % 2) * 1.4, rng.normal(scale=0.3, size=nat)], -1)
    out["chain150"] = (zz, xyz, zero_sum_charges(zz, 42), TPSSH, {})

    # tight custom cutoffs on a blob: many open triples / cut pairs
    zz, xyz, q = orc.organic_blob(40, np.random.default_rng(51))
    out["tight_cutoffs"] = (zz, xyz, q, dict(PBE0, s9=1.3), dict(disp2=9.0, disp3=6.5))

    # strongly charged atoms incl. q + Z <= 0 branch of zeta (model/base.py:331-335)
    zz, xyz, q = orc.organic_blob(12, np.random.default_rng(61))
    q = q * 0 + np.linspace(-1.2, 1.2, 12)
    out["big_charges"] = (zz, xyz, q, PBE0, {})
    return out


def run_reference(numbers, positions, q, param, kw, model):
    n = torch.from_numpy(np.asarray(numbers)).to(torch.int64)
    p = torch.from_numpy(np.asarray(positions, dtype=np.float64)).clone().requires_grad_(True)
    qq = torch.from_numpy(np.asarray(q, dtype=np.float64))
    par = {k: torch.tensor(v, dtype=F64) for k, v in param.items()}
    cut = ref.Cutoff(dtype=F64, **kw) if kw else None
    e = ref.dftd4(n, p, torch.zeros(n.shape[:-1], dtype=F64), par, q=qq, model=model, cutoff=cut)
    (g,) = torch.autograd.grad(e.sum(), p)
    return e.detach(), g


def run_oracle(numbers, positions, q, param, kw, model):
    n = torch.from_numpy(np.asarray(numbers)).to(torch.int64)
    p = torch.from_numpy(np.asarray(positions, dtype=np.float64))
    qq = torch.from_numpy(np.asarray(q, dtype=np.float64))
    return orc.energy_and_gradient(n, p, param, qq, model=model, **kw)


def main() -> None:
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    worst_e = worst_g = 0.0
    for name, (numbers, positions, q, param, kw) in cases().items():
        store = {
            "numbers": np.asarray(numbers, dtype=np.int64),
            "positions": np.asarray(positions, dtype=np.float64),
            "q": np.asarray(q, dtype=np.float64),
            "param_keys": np.array(sorted(param)),
            "param_vals": np.array([param[k] for k in sorted(param)], dtype=np.float64),
            "cutoff_keys": np.array(sorted(kw)),
            "cutoff_vals": np.array([kw[k] for k in sorted(kw)], dtype=np.float64),
        }
        for model in ("d4", "d4s"):
            e, g = run_reference(numbers, positions, q, param, kw, model)
            eo, go = run_oracle(numbers, positions, q, param, kw, model)
            scale = e.abs().max().clamp(min=1e-30)
            de = ((e - eo).abs().max() / scale).item()
            dg = (g - go).abs().max().item()
            worst_e, worst_g = max(worst_e, de), max(worst_g, dg)
            print(f"{name:20s} {model:4s} E={e.sum().item(): .12e} oracle-ref: dE/max|E|={de:.1e} dG={dg:.1e}")
            assert de < 1e-12 and dg < 1e-13, (name, model, de, dg)
            store[f"energy_{model}"] = e.numpy()
            store[f"grad_{model}"] = g.numpy()

        # properties through the reference's own model classes (q explicit)
        n = torch.from_numpy(store["numbers"])
        p = torch.from_numpy(store["positions"])
        qq = torch.from_numpy(store["q"])
        cn = ref_cn_d4(n, p)
        mdl = ref.D4Model(n, dtype=F64)
        w = mdl.weight_references(cn, qq)
        c6 = mdl.get_atomic_c6(w)
        alpha = mdl.get_polarizabilities(w)
        ocn, _, oc6, oalpha = orc.get_properties(n, p, qq)
        assert torch.allclose(cn, ocn, rtol=1e-13, atol=1e-15)
        assert torch.allclose(c6, oc6, rtol=1e-12, atol=1e-14)
        assert torch.allclose(alpha, oalpha, rtol=1e-12, atol=1e-14)
        store["cn"], store["c6"], store["alpha"] = cn.numpy(), c6.numpy(), alpha.numpy()
        np.savez_compressed(OUT / f"{name}.npz", **store)
    print(f"worst oracle-vs-reference: energy rel {worst_e:.2e}, gradient abs {worst_g:.2e}")


if __name__ == "__main__":
    main()
