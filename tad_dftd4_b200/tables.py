"""
Host-side table compiler for the CUDA kernels.

The reference rebuilds, on every ``dftd4`` call, the reference polarizabilities
``alpha[i,a,w]`` and the ``(N, N, 7, 7)`` reference-C6 tensor from element
tables gathered by ``numbers`` (``/root/reference/src/tad_dftd4/model/base.py:
367-431``, ``utils.py:33-94``).  Both are pure functions of the atomic number,
so here they are compiled ONCE per ``(ga, gc)`` into flat per-element arrays
that live in HBM/L2 for the lifetime of the process:

``alpha_w[Z, a, w] = sqrt(3/pi * trapz_w) * max(0, ascale*(refalpha - scount*sec*zeta))``

so that the pair coefficient becomes a 23-term dot product of two per-atom
vectors, ``C6_ij = sum_w A_i[w] A_j[w]`` with ``A_i[w] = sum_a W_ia
alpha_w[Z_i, a, w]`` -- algebraically identical to the reference's
``einsum('ijab,ia,jb->ij', rc6, W, W)`` (``model/d4.py:285-289``) but without
any per-pair table.

Everything here is plain numpy float64; the blob is handed to the C-ABI
(``d4b200_tables_create``) as one double array + one int32 array.
"""

from __future__ import annotations

from functools import lru_cache
from pathlib import Path

import numpy as np

from .data import elements as _el

__all__ = ["ElementTables", "build_tables", "NELEM", "NREF", "NFREQ"]

_DATA = Path(__file__).resolve().parent / "data"

NELEM = 104  # reference tables cover Z = 0 (dummy) .. 103
NREF = 7
NFREQ = 23

# defaults.py:41-51 of the reference
KCN = 7.5
K4 = 4.10451
K5 = 19.08857
K6 = 2 * 11.28174**2

# utils.py:52-80: trapezoid weights of the Casimir-Polder quadrature
CP_WEIGHTS = np.array(
    [
        2.4999500000000000e-002, 4.9999500000000000e-002, 7.5000000000000010e-002,
        0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.15, 0.2, 0.2, 0.2, 0.2, 0.35,
        0.5, 0.75, 1.0, 1.75, 2.5, 1.25,
    ]
)  # fmt: skip
THOPI = 3.0 / 3.141592653589793238462643383279502884197


@lru_cache(maxsize=None)
def _raw() -> dict[str, np.ndarray]:
    with np.load(_DATA / "d4_reference.npz") as f:
        return {k: f[k] for k in f.files}


def _zeta(gam, qref, qmod, ga):
    """model/base.py:326-335 (float64: eps = 2.22e-16)."""
    eps = np.finfo(np.float64).eps
    with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
        scale = np.exp(gam * (1.0 - qref / (qmod - eps)))
        return np.where(qmod > 0.0, np.exp(ga * (1.0 - scale)), np.exp(ga))


class ElementTables:
    """Flat per-element tables (numpy, float64 / int32)."""

    # order of the sections inside the double blob (see csrc/d4b200_tables.cuh)
    F64_LAYOUT = (
        ("rcov", NELEM),
        ("r4r2", NELEM),
        ("sqrt_r4r2", NELEM),  # 3^(1/4) * sqrt(r4r2): R0_ij = a1 * s_i * s_j + a2
        ("gamgc", NELEM),
        ("zeff", NELEM),
        ("refcn", NELEM * NREF),
        ("refq", NELEM * NREF),  # clsq + zeff
        ("zeta0", NELEM * NREF),  # zeta at q = 0 (ATM flavour), 0 where refc == 0
        ("alpha0", NELEM * NREF),  # static reference polarizability alpha(i w=0)
        ("den", NELEM * NELEM),  # CN electronegativity factor
        ("alpha_w", NELEM * NREF * NFREQ),
        ("wfpair", NELEM * NELEM),
        ("rc6", NELEM * NELEM * NREF * NREF),  # reference C6 per element pair (D4S)
    )
    I32_LAYOUT = (
        ("refc", NELEM * NREF),
        ("maxcn_ref", NELEM),  # index of the largest reference CN (NaN fallback)
    )

    def __init__(self, ga: float, gc: float, ref_charges: str = "eeq", c9_frequency: int | None = None):
        raw = _raw()
        self.ga, self.gc = float(ga), float(gc)
        # reference charges of the model (model/base.py:388-399, model/d4.py:142-149): EEQ (default) or GFN2-xTB
        if ref_charges not in ("eeq", "gfn2"):
            raise ValueError(f"Unknown reference charges: {ref_charges}")
        self.ref_charges = ref_charges
        ref_q, ref_h = raw["clsq"], raw["clsh"]
        if ref_charges == "gfn2":  # tabulated for Z <= 86 (reference/d4/charge_gfn2.py); zero rows beyond
            ref_q, ref_h = np.zeros_like(ref_q), np.zeros_like(ref_h)
            ref_q[: raw["gfn2_refq"].shape[0]] = raw["gfn2_refq"]
            ref_h[: raw["gfn2_refh"].shape[0]] = raw["gfn2_refh"]
        z = np.arange(NELEM)
        gam = np.asarray(_el.GAM, dtype=np.float64)
        zeff = np.asarray(_el.ZEFF, dtype=np.float64)
        en = np.asarray(_el.PAULING, dtype=np.float64)[:NELEM]

        self.rcov = np.asarray(_el.COV_2009, dtype=np.float64)[:NELEM] * _el.AA2AU * 4.0 / 3.0
        self.r4r2 = raw["r4r2"][:NELEM].copy()
        self.sqrt_r4r2 = 3.0**0.25 * np.sqrt(self.r4r2)
        self.gamgc = gam[:NELEM] * gc
        self.zeff = zeff[:NELEM].copy()
        self.refcn = raw["refcovcn"].copy()
        self.refc = raw["refc"].astype(np.int32)
        self.refq = ref_q + self.zeff[:, None]
        mask = self.refc > 0
        self.zeta0 = np.where(
            mask, _zeta(self.gamgc[:, None], self.refq, self.zeff[:, None], ga), 0.0
        )
        self.maxcn_ref = np.argmax(self.refcn, axis=1).astype(np.int32)

        # CN: den_ij = k4 * exp(-(|en_i - en_j| + k5)^2 / k6)
        self.den = K4 * np.exp(-((np.abs(en[:, None] - en[None, :]) + K5) ** 2) / K6)

        # reference polarizabilities, model/base.py:379-418
        refsys = raw["refsys"].astype(np.int64)
        zs = zeff[refsys]
        gs = gam[refsys] * gc
        zsec = np.where(refsys > 0, _zeta(gs, zs, ref_h + zs, ga), 0.0)
        sec = raw["secscale"][refsys] * raw["secalpha"][refsys] * zsec[..., None]
        alpha = raw["refascale"][..., None] * (raw["refalpha"] - raw["refscount"][..., None] * sec)
        self.alpha = np.where(alpha > 0.0, alpha, 0.0)  # (104, 7, 23)
        self.alpha0 = self.alpha[..., 0].copy()
        self.alpha_w = self.alpha * np.sqrt(THOPI * CP_WEIGHTS)[None, None, :]
        self.c9_frequency = c9_frequency
        if c9_frequency is not None:
            # One Casimir-Polder node of the EXACT C9 (dispersion/threebody.py:276-302, utils.py:155-212):
            #   C9_ijk = (3/pi) sum_w t_w a_i(w) a_j(w) a_k(w)
            # and, all polarizabilities being >= 0, each node is the kernels' own approximate form
            # sqrt(c_ij c_jk c_ik) with the pair quantity c_ij = b_i b_j, b_i = ((3/pi) t_w)^(1/3) a_i(w).  The
            # kernels build c_ij as the dot product of the weighted-polarizability vectors, so a table that
            # keeps only node w, rescaled from sqrt((3/pi) t_w) to the cube root, makes an ATM-only launch
            # return exactly that node's contribution, gradients and all (dispersion.D4ATMExact sums the 23).
            w = int(c9_frequency)
            if not 0 <= w < NFREQ:
                raise ValueError(f"Casimir-Polder node {w} outside 0..{NFREQ - 1}")
            node = np.zeros_like(self.alpha_w)
            node[..., w] = self.alpha_w[..., w] * (THOPI * CP_WEIGHTS[w]) ** (-1.0 / 6.0)
            self.alpha_w = node
        self.wfpair = raw["wfpair"][:NELEM, :NELEM].copy()
        # rc6[Za, Zb, a, b] = (3/pi) sum_w w_w alpha[Za,a,w] alpha[Zb,b,w]  (utils.py:91-94)
        flat = self.alpha_w.reshape(NELEM * NREF, NFREQ)
        self.rc6 = (flat @ flat.T).reshape(NELEM, NREF, NELEM, NREF).transpose(0, 2, 1, 3).copy()

    def f64_blob(self) -> np.ndarray:
        parts = []
        for name, size in self.F64_LAYOUT:
            a = np.ascontiguousarray(getattr(self, name), dtype=np.float64).reshape(-1)
            assert a.size == size, (name, a.size, size)
            parts.append(a)
        return np.concatenate(parts)

    def i32_blob(self) -> np.ndarray:
        parts = []
        for name, size in self.I32_LAYOUT:
            a = np.ascontiguousarray(getattr(self, name), dtype=np.int32).reshape(-1)
            assert a.size == size, (name, a.size, size)
            parts.append(a)
        return np.concatenate(parts)


@lru_cache(maxsize=64)
def build_tables(ga: float = 3.0, gc: float = 2.0, ref_charges: str = "eeq",
                 c9_frequency: int | None = None) -> ElementTables:  # fmt: skip
    return ElementTables(ga, gc, ref_charges, c9_frequency)
