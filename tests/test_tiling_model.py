"""Index logic of the register-tiled triple loops (csrc/d4b200_small.cuh: energy_triples_tiled,
grad_sweep_tiled) replayed on the CPU: every triple exactly once (energy), every (owner pair, third
atom) exactly once (gradient), for even / odd sizes, any number of warps and every lane-sharing mode."""
from __future__ import annotations

from collections import Counter
from itertools import combinations

import pytest


def pair_lookup(p):
    hi = 1
    while (hi + 1) * hi // 2 <= p:
        hi += 1
    return hi, p - hi * (hi - 1) // 2


def energy_triples(n, nw):
    seen = Counter()
    n_i = n >> 1
    nblk = n_i * (n_i - 1) // 2
    nchunks = (nblk + 31) >> 5
    for rnd in range((nchunks + nw - 1) // nw if nchunks else 0):
        for warp in range(nw):
            chunk = rnd * nw + (nw - 1 - warp if rnd & 1 else warp)
            if chunk >= nchunks:
                continue
            lanes = []
            for lane in range(32):
                item = chunk * 32 + lane
                valid = item < nblk
                J, m = pair_lookup(item if valid else nblk - 1)
                lanes.append((valid, J, m))
            jfirst = lanes[0][1]
            for valid, J, m in lanes:
                j0, j1, k0, k1 = 2 * J, 2 * J + 1, m, m + J
                assert k1 < j0
                i0 = 2 * jfirst + 2
                while i0 < n:
                    for u in range(4):
                        i = i0 + u
                        if i < n and valid and i > j1:
                            for j in (j0, j1):
                                for k in (k0, k1):
                                    seen[(i, j, k)] += 1
                    i0 += 4
                if valid:
                    seen[(j1, j0, k0)] += 1
                    seen[(j1, j0, k1)] += 1
    nd = n >> 1
    c = 0
    while c * 32 < nd:
        for lane in range(32):
            d = c * 32 + lane
            if d < nd:
                j, k = 2 * d + 1, 2 * d
                i0 = 64 * c + 2
                while i0 < n:
                    for u in range(4):
                        i = i0 + u
                        if i < n and i > j:
                            seen[(i, j, k)] += 1
                    i0 += 4
        c += 1
    return seen


@pytest.mark.parametrize("n", [4, 5, 6, 7, 20, 33, 48, 60, 61, 64, 97, 120])
@pytest.mark.parametrize("nw", [4, 8, 16])
def test_energy_tiling_visits_every_triple_once(n, nw):
    seen = energy_triples(n, nw)
    want = {(i, j, k) for k, j, i in combinations(range(n), 3)}
    assert set(seen) == want and all(v == 1 for v in seen.values())


def gradient_visits(n, nt):
    seen = Counter()
    n_i = (n + 1) >> 1
    nblk = n_i * (n_i - 1) // 2
    base = 0
    while base < nblk:
        left = nblk - base
        parts = 1 if left * 2 > nt else (2 if left * 4 > nt else 4)
        per = 32 // parts
        for warp in range(nt // 32):
            if warp * per >= left:
                continue
            for lane in range(32):
                sub = lane // per
                item = base + warp * per + (lane - sub * per)
                if item >= nblk:
                    continue
                I, m = pair_lookup(item)
                i0, k0, k1 = 2 * I, m, m + I
                has1 = i0 + 1 < n
                i1 = i0 + 1 if has1 else i0
                length = (n + parts - 1) // parts
                for j in range(sub * length, min(n, sub * length + length)):
                    for a, ia in enumerate((i0, i1)):
                        if a == 1 and not has1:
                            continue
                        for k in (k0, k1):
                            if j != ia and j != k:
                                seen[(ia, k, j)] += 1
        base += nt
    nd = n >> 1
    len8 = (n + 7) >> 3
    for d in range(nd):
        jj, kk = 2 * d + 1, 2 * d
        for sub in range(8):
            for j in range(sub * len8, min(n, (sub + 1) * len8)):
                if j != jj and j != kk:
                    seen[(jj, kk, j)] += 1
    return seen


@pytest.mark.parametrize("n", [4, 5, 7, 20, 33, 64, 65, 80, 99, 100])
@pytest.mark.parametrize("nt", [128, 256, 512])
def test_gradient_tiling_visits_every_owner_pair_and_third_atom_once(n, nt):
    seen = gradient_visits(n, nt)
    want = {(i, k, j) for k, i in combinations(range(n), 2) for j in range(n) if j not in (i, k)}
    assert set(seen) == want and all(v == 1 for v in seen.values())
