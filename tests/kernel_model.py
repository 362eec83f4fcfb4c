"""
Numpy model of the ALGEBRA the CUDA kernels use (not of their parallel
structure): per-atom weighted-polarizability vectors instead of the 7x7
reference contraction, the factorised per-pair ATM stash and the analytic
gradient of SURVEY.md Appendix D.  ``tests/test_kernel_model.py`` checks it
against the oracle so that formula errors are found on the CPU before a GPU
run; DESIGN.md points here for the derivation.

Test infrastructure only.
"""

from __future__ import annotations

import numpy as np
from scipy.special import erfc

from tad_dftd4_b200.tables import KCN, NREF, build_tables


def _weights(tab, z, cn, q, wf, ga):
    """Gaussian weights (max-shifted evaluation), zeta scaling and their
    derivatives for one atom list.  Returns gw, dgw/dcn, zeta, dzeta/dq, zeta0."""
    n = len(z)
    refc = tab.refc[z]
    valid = refc > 0
    d = cn[:, None] - tab.refcn[z]
    arg = wf * d * d
    shift = np.min(np.where(valid, arg, np.inf), axis=1, keepdims=True)
    shift = np.where(np.isfinite(shift), shift, 0.0)
    s = np.zeros((n, NREF))
    ds = np.zeros((n, NREF))
    for k in (1, 2, 3):
        use = valid & (refc >= k)
        e = np.where(use, np.exp(-(k * arg - shift)), 0.0)
        s += e
        ds += -2.0 * k * wf * d * e
    norm = s.sum(1, keepdims=True)
    norm = np.where(norm > 0, norm, 1.0)
    gw = s / norm
    dgw = (ds - gw * ds.sum(1, keepdims=True)) / norm

    eps = np.finfo(np.float64).eps
    gam = tab.gamgc[z][:, None]
    qref = tab.refq[z]
    qmod = (q + tab.zeff[z])[:, None]
    with np.errstate(all="ignore"):
        scale = np.exp(gam * (1.0 - qref / (qmod - eps)))
        zeta = np.where(qmod > 0, np.exp(ga * (1.0 - scale)), np.exp(ga))
        dzeta = np.where(qmod > 0, -ga * gam * scale * zeta * qref / (qmod - eps) ** 2, 0.0)
    zeta = np.where(valid, zeta, 0.0)
    dzeta = np.where(valid, dzeta, 0.0)
    return gw, dgw, zeta, dzeta, tab.zeta0[z]


def energy_and_gradient(numbers, positions, q, param, g=None, disp2=60.0, disp3=40.0,
                        cn_cut=30.0, wf=6.0, ga=3.0, gc=2.0):  # fmt: skip
    """Single structure (1-D numbers, no padding).  Returns per-atom energy,
    dL/dpositions and dL/dq for L = sum_i g_i E_i."""
    tab = build_tables(ga, gc)
    z = np.asarray(numbers)
    x = np.asarray(positions, dtype=np.float64)
    n = len(z)
    g = np.ones(n) if g is None else np.asarray(g, dtype=np.float64)
    s6, s8, s9 = param.get("s6", 1.0), param.get("s8", 1.0), param.get("s9", 1.0)
    s10 = param.get("s10", 0.0) if "s10" in param else 0.0
    a1, a2, alp = param["a1"], param["a2"], param.get("alp", 16.0)

    dx = x[:, None, :] - x[None, :, :]
    r2 = (dx * dx).sum(-1)
    off = ~np.eye(n, dtype=bool)
    r = np.sqrt(np.where(off, r2, 1.0))

    # ---- CN and its radial derivative ------------------------------------
    r0 = tab.rcov[z][:, None] + tab.rcov[z][None, :]
    den = tab.den[z][:, z]
    xx = KCN * (r / r0 - 1.0)
    incn = off & (r <= cn_cut)
    cn = np.where(incn, den * 0.5 * erfc(xx), 0.0).sum(1)
    dcn_dr = np.where(incn, -den * KCN / (r0 * np.sqrt(np.pi)) * np.exp(-xx * xx), 0.0)

    # ---- weights and per-atom polarizability vectors ----------------------
    gw, dgw, zeta, dzeta, zeta0 = _weights(tab, z, cn, np.asarray(q, dtype=np.float64), wf, ga)
    aw = tab.alpha_w[z]  # (n, 7, 23)
    Aq = np.einsum("ia,iaw->iw", zeta * gw, aw)
    A0 = np.einsum("ia,iaw->iw", zeta0 * gw, aw)
    c6q = Aq @ Aq.T
    c60 = A0 @ A0.T

    # ---- two-body ---------------------------------------------------------
    sq = tab.sqrt_r4r2[z]
    R0 = a1 * sq[:, None] * sq[None, :] + a2
    qq = 3.0 * tab.r4r2[z][:, None] * tab.r4r2[z][None, :]
    in2 = off & (r <= disp2)
    t6 = 1.0 / (r**6 + R0**6)
    t8 = 1.0 / (r**8 + R0**8)
    t10 = 1.0 / (r**10 + R0**10)
    k10 = s10 * 49.0 / 40.0 * qq * qq
    F = np.where(in2, s6 * t6 + s8 * qq * t8 + k10 * t10, 0.0)
    dF = np.where(
        in2,
        -(6 * s6 * r**5 * t6**2 + 8 * s8 * qq * r**7 * t8**2 + 10 * k10 * r**9 * t10**2),
        0.0,
    )
    e2 = -0.5 * (c6q * F).sum(1)
    G2 = -0.5 * (g[:, None] + g[None, :])

    # ---- ATM: factorised per-pair stash ----------------------------------
    fac = np.cbrt(s9 / 6.0)
    P = np.where(off, fac * np.sqrt(np.abs(c60)) / (r2 * r), 0.0)
    ap = alp / 3.0
    u = np.where(off, (R0 / r) ** ap, 0.0)
    cflag = (off & (r <= disp3)).astype(np.float64)
    a = np.where(off, r2, 1.0)

    e3 = np.zeros(n)
    Gam = np.zeros((n, n))  # sum_triples W e (symmetric, later / (2 c60))
    D = np.zeros((n, n))  # dL3 / d r2_p
    for i in range(n):
        for j in range(i):
            for k in range(j):
                aij, aik, ajk = a[i, j], a[i, k], a[j, k]
                cij, cik, cjk = cflag[i, j], cflag[i, k], cflag[j, k]
                mi, mj, mk = cjk * (cij + cik), cik * (cij + cjk), cij * (cik + cjk)
                if mi + mj + mk == 0:
                    continue
                X, Y, Zz = aij + ajk - aik, aij - ajk + aik, -aij + ajk + aik
                s = X * Y * Zz
                Q = 1.0 / (aij * aik * ajk)
                PS = P[i, j] * P[i, k] * P[j, k]
                t = u[i, j] * u[i, k] * u[j, k]
                f = 1.0 / (1.0 + 6.0 * t)
                ang = 0.375 * s * Q + 1.0
                e = ang * PS * f  # = e_ijk / 6
                e3[i] += mi * e
                e3[j] += mj * e
                e3[k] += mk * e
                W = g[i] * mi + g[j] * mj + g[k] * mk
                psf = PS * f
                common = e * (-2.5 + 3.0 * ap * f * t) + psf
                # s = X Y Z with X=aij+ajk-aik, Y=aij-ajk+aik, Z=-aij+ajk+aik
                ds_ij = Y * Zz + X * Zz - X * Y
                ds_jk = Y * Zz - X * Zz + X * Y
                ds_ik = -Y * Zz + X * Zz + X * Y
                for p, ds, av in (((i, j), ds_ij, aij), ((j, k), ds_jk, ajk), ((i, k), ds_ik, aik)):
                    de = common / av + 0.375 * psf * Q * ds
                    D[p] += W * de
                    D[p[::-1]] += W * de
                    Gam[p] += W * e
                    Gam[p[::-1]] += W * e
    with np.errstate(all="ignore"):
        Gam = np.where(off & (c60 != 0), Gam / (2.0 * c60), 0.0)

    # ---- chain rule through C6 -> weights -> (cn, q) ----------------------
    Bq = (G2 * F) @ Aq  # dL/dAq
    B0 = Gam @ A0  # dL/dA0
    projq = np.einsum("iaw,iw->ia", aw, Bq)
    proj0 = np.einsum("iaw,iw->ia", aw, B0)
    dL_dcn = (zeta * dgw * projq).sum(1) + (zeta0 * dgw * proj0).sum(1)
    dL_dq = (dzeta * gw * projq).sum(1)

    # ---- forces: everything is a pair coefficient times (R_i - R_j) -------
    fc = np.where(off, G2 * c6q * dF / r + 2.0 * D + (dL_dcn[:, None] + dL_dcn[None, :]) * dcn_dr / r, 0.0)
    grad = (fc[:, :, None] * dx).sum(1)
    return e2 + e3, grad, dL_dq, cn


def param_gradient(numbers, positions, q, param, g=None, disp2=60.0, disp3=40.0, cn_cut=30.0,
                   wf=6.0, ga=3.0, gc=2.0):  # fmt: skip
    """dL/d(s6, s8, s9, s10, a1, a2, alp) for L = sum_i g_i E_i, the algebra of
    ``csrc/d4b200_param.cu`` (single structure, D4 model): the energy is linear in the
    scaling factors; a1/a2 enter through R0 = a1 sqrt(3 r4r2_i r4r2_j) + a2 in the rational
    damping and, together with alp, through u = (R0/r)^(alp/3) in the ATM damping
    f = 1/(1 + 6 u_ij u_ik u_jk), so d e/d ln u_p = -6 t f e for each pair of a triple."""
    tab = build_tables(ga, gc)
    z = np.asarray(numbers)
    x = np.asarray(positions, dtype=np.float64)
    n = len(z)
    g = np.ones(n) if g is None else np.asarray(g, dtype=np.float64)
    s6, s8, s9 = param.get("s6", 1.0), param.get("s8", 1.0), param.get("s9", 1.0)
    has_s10 = "s10" in param
    s10 = param.get("s10", 0.0) if has_s10 else 0.0
    a1, a2, alp = param["a1"], param["a2"], param.get("alp", 16.0)

    dx = x[:, None, :] - x[None, :, :]
    r2 = (dx * dx).sum(-1)
    off = ~np.eye(n, dtype=bool)
    r = np.sqrt(np.where(off, r2, 1.0))
    r0 = tab.rcov[z][:, None] + tab.rcov[z][None, :]
    den = tab.den[z][:, z]
    cn = np.where(off & (r <= cn_cut), den * 0.5 * erfc(KCN * (r / r0 - 1.0)), 0.0).sum(1)
    gw, _, zeta, _, zeta0 = _weights(tab, z, cn, np.asarray(q, dtype=np.float64), wf, ga)
    aw = tab.alpha_w[z]
    Aq = np.einsum("ia,iaw->iw", zeta * gw, aw)
    A0 = np.einsum("ia,iaw->iw", zeta0 * gw, aw)
    c6q, c60 = Aq @ Aq.T, A0 @ A0.T

    sq = tab.sqrt_r4r2[z]
    ss = sq[:, None] * sq[None, :]
    R0 = a1 * ss + a2
    qq = ss * ss
    in2 = off & (r <= disp2)
    t6, t8, t10 = 1.0 / (r**6 + R0**6), 1.0 / (r**8 + R0**8), 1.0 / (r**10 + R0**10)
    k10 = 49.0 / 40.0 * qq * qq
    G2c = np.where(in2, -0.25 * (g[:, None] + g[None, :]) * c6q, 0.0)  # every pair appears twice
    dFdR0 = -(6 * s6 * R0**5 * t6**2 + 8 * s8 * qq * R0**7 * t8**2 + 10 * s10 * k10 * R0**9 * t10**2)
    out = np.zeros(7)
    out[0] = (G2c * t6).sum()
    out[1] = (G2c * qq * t8).sum()
    out[3] = (G2c * k10 * t10).sum() if has_s10 else 0.0
    out[4] = (G2c * dFdR0 * ss).sum()
    out[5] = (G2c * dFdR0).sum()

    Pt = np.where(off, np.sqrt(np.abs(c60)) / (r2 * r2 * r), 0.0)
    u = np.where(off, (R0 / r) ** (alp / 3.0), 0.0)
    lg = np.log(np.where(off, R0 / r, 1.0))
    cflag = (off & (r <= disp3)).astype(np.float64)
    for i in range(n):
        for j in range(i):
            for k in range(j):
                a, b, c = r2[i, j], r2[j, k], r2[i, k]
                cij, cik, cjk = cflag[i, j], cflag[i, k], cflag[j, k]
                W = g[i] * cjk * (cij + cik) + g[j] * cik * (cij + cjk) + g[k] * cij * (cik + cjk)
                if W == 0.0:
                    continue
                s = (b * b - (a - c) ** 2) * (a + c - b)
                t = u[i, j] * u[i, k] * u[j, k]
                f = 1.0 / (1.0 + 6.0 * t)
                e = W * (0.375 * s + a * b * c) * Pt[i, j] * Pt[i, k] * Pt[j, k] * f / 6.0
                out[2] += e
                h = -6.0 * t * f * e * s9
                out[4] += h * alp / 3.0 * (ss[i, j] / R0[i, j] + ss[i, k] / R0[i, k] + ss[j, k] / R0[j, k])
                out[5] += h * alp / 3.0 * (1.0 / R0[i, j] + 1.0 / R0[i, k] + 1.0 / R0[j, k])
                out[6] += h * (lg[i, j] + lg[i, k] + lg[j, k]) / 3.0
    return out


def grad_visit_reference(a, b, c, Pij, Pik, Pjk, uij, uik, ujk, alp):
    """One owner-pair visit, textbook form: e' and d e'/d b for the owner pair (j,k) with
    squared distances a = r_ij^2, b = r_jk^2, c = r_ik^2 (threebody.py:113-160 per pair)."""
    s = (a + b - c) * (a - b + c) * (b + c - a)
    dsdb = (a - b + c) * (b + c - a) - (a + b - c) * (b + c - a) + (a + b - c) * (a - b + c)
    t = uij * uik * ujk
    f = 1.0 / (1.0 + 6.0 * t)
    pf = Pij * Pik * Pjk * f
    abc = a * b * c
    e = pf * (0.375 * s + abc)
    de = (e * (alp * f * t - 2.5) + pf * abc) / b + 0.375 * pf * dsdb
    return e, de


def grad_visit_kernel(a, b, c, Pij, Pik, Pjk, uij, uik, ujk, alp):
    """The same visit in the form ``grad_visit<T, false, true>`` of csrc/d4b200_small.cuh evaluates
    it (21 FP64 instructions): owner-pair invariants folded into the damping denominator,
    f t = (1 - f)/6, the pair factor applied in the accumulating FMAs, 1/b and 0.375 applied
    once per owner pair.  Returns the three accumulator increments and the assembled (e', de')."""
    iP = 1.0 / Pjk
    sPu, kA = 6.0 * ujk * iP, alp / 6.0
    kAi, kB = kA * iP, kA - 2.5
    t1, Y = a - c, (a + c) - b
    XZ = b * b - t1 * t1
    s = XZ * Y
    dsdb = 2.0 * b * Y - XZ
    abc = (a * c) * b
    fp = 1.0 / (iP + sPu * (uij * uik))  # = P'_jk f
    pf = (Pij * Pik) * fp
    wa = 0.375 * s + abc
    inner = wa * (kB - kAi * fp) + abc
    dG, dC, dS = pf * wa, pf * inner, pf * dsdb
    return (dG, dC, dS), (dG, dC / b + 0.375 * dS)
