"""
numpy model of the device EEQ kernels (``tad_dftd4_b200/csrc/d4b200_eeq.cu``): the same
formulas in the same order -- compacted atoms, direct-difference distances, near-pair
counting, bordered system eliminated WITHOUT pivoting (A is a Gram matrix of Gaussian
charges + hardness: positive definite for physical geometries; the border pivot is
-1^T A^-1 1), and the analytic vector-Jacobian product

    dL/dx = mu^T (d rhs/dx - dM/dx x),   M mu = (dL/dq, 0)

Checked on the CPU against the dense oracle + autograd (``tests/test_eeq_model.py``).
"""
from __future__ import annotations

import math

import numpy as np

SQRT2PI = math.sqrt(2.0 / math.pi)
TWO_SQRTPI = 2.0 / math.sqrt(math.pi)
KCN = 7.5
CN_MAX = 8.0
ZSKIP = 6.0  # erfc(6)/2 = 1e-17: below one ulp of any coordination number that matters
EPS = np.finfo(np.float64).eps

try:  # scipy is test-only
    from scipy.special import erf, erfc
except Exception:  # pragma: no cover
    erf = np.vectorize(math.erf)
    erfc = np.vectorize(math.erfc)


def _setup(z, xyz, par):
    idx = np.nonzero(z)[0]
    zz = z[idx]
    x = xyz[idx]
    return idx, zz, x, par["chi"][zz], par["eta"][zz], par["kcn"][zz], par["rad"][zz], par["rcov"][zz]


def _count(x, rcov, cutoff):
    n = len(x)
    cn = np.zeros(n)
    for i in range(n):
        d = x[i] - x
        r2 = (d * d).sum(-1)
        r = np.sqrt(r2)
        zarg = KCN * (r / (rcov[i] + rcov) - 1.0)
        ok = (np.arange(n) != i) & (r2 <= cutoff * cutoff) & (zarg < ZSKIP)
        cn[i] = (0.5 * erfc(zarg[ok])).sum()
    return cn


def _cut(cn):
    return math.log1p(math.exp(CN_MAX)) - np.log1p(np.exp(CN_MAX - cn))


def _matrix(x, eta, rad):
    n = len(x)
    m = np.zeros((n + 1, n + 1))
    for i in range(n):
        for j in range(i):
            r = math.sqrt(((x[i] - x[j]) ** 2).sum())
            g = 1.0 / math.sqrt(rad[i] ** 2 + rad[j] ** 2)
            m[i, j] = m[j, i] = math.erf(g * r) / r
        m[i, i] = eta[i] + SQRT2PI / rad[i]
        m[i, n] = m[n, i] = 1.0
    return m


def _solve_nopivot(m, b):
    """Forward elimination of [M | b] without pivoting, then back substitution."""
    a = np.concatenate([m, b[:, None]], axis=1)
    k = a.shape[0]
    for c in range(k):
        f = a[c + 1:, c] / a[c, c]
        a[c + 1:, c + 1:] -= f[:, None] * a[c, c + 1:][None, :]
    xs = np.zeros(k)
    for c in range(k - 1, -1, -1):
        xs[c] = a[c, k] / a[c, c]
        a[:c, k] -= a[:c, c] * xs[c]
    return xs


def eeq_forward(z, xyz, charge, par, cutoff=25.0):
    idx, zz, x, chi, eta, kap, rad, rcov = _setup(z, xyz, par)
    n = len(idx)
    q = np.zeros(len(z))
    if n == 0:
        return q
    cn = _cut(_count(x, rcov, cutoff))
    rhs = np.concatenate([-chi + kap * np.sqrt(np.maximum(cn, EPS)), [charge]])
    sol = _solve_nopivot(_matrix(x, eta, rad), rhs)
    q[idx] = sol[:n]
    return q


def eeq_backward(z, xyz, q, gq, par, cutoff=25.0):
    idx, zz, x, chi, eta, kap, rad, rcov = _setup(z, xyz, par)
    n = len(idx)
    grad = np.zeros_like(xyz)
    if n == 0:
        return grad
    raw = _count(x, rcov, cutoff)
    cn = _cut(raw)
    mu = _solve_nopivot(_matrix(x, eta, rad), np.concatenate([gq[idx], [0.0]]))[:n]
    qq = q[idx]
    # dL/dcn_raw
    big = np.where(cn > EPS, mu * kap * 0.5 / np.sqrt(np.maximum(cn, EPS)), 0.0) / (1.0 + np.exp(raw - CN_MAX))
    g = np.zeros((n, 3))
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            d = x[i] - x[j]
            r2 = (d * d).sum()
            r = math.sqrt(r2)
            gam = 1.0 / math.sqrt(rad[i] ** 2 + rad[j] ** 2)
            da = TWO_SQRTPI * gam * math.exp(-gam * gam * r2) / r - math.erf(gam * r) / r2
            w = -(mu[i] * qq[j] + mu[j] * qq[i]) * da
            r0 = rcov[i] + rcov[j]
            zarg = KCN * (r / r0 - 1.0)
            if r2 <= cutoff * cutoff and zarg * zarg < 40.0:
                w += (big[i] + big[j]) * (-KCN / (r0 * math.sqrt(math.pi))) * math.exp(-zarg * zarg)
            g[i] += (w / r) * d
    grad[idx] = g
    return grad
