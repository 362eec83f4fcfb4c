"""
Synthetic inputs of the benchmark workloads: the organic blob
generator, padded batches of blobs, the 20 001-atom water cluster (C4) and the
``examples/single.py`` molecule (C1).

Pure numpy input generation -- no D4 arithmetic.  Shared by ``bench.py`` (both arms),
the tests and the oracle's golden-vector scripts (``oracle/d4_oracle.py`` re-exports the
organic generators under their historical names).
"""

from __future__ import annotations

import numpy as np


def _torch():
    import torch  # lazily: the generator's worker processes only need numpy

    return torch

# --------------------------------------------------------------------------
# organic blobs (C2, C3, C5)
# --------------------------------------------------------------------------
_ORGANIC_Z = np.array([1, 6, 7, 8, 16, 9])
_ORGANIC_P = np.array([0.48, 0.32, 0.07, 0.10, 0.02, 0.01])
_VALENCE = {1: 1, 6: 4, 7: 3, 8: 2, 16: 2, 9: 1}


def organic_blob(nat: int, rng: np.random.Generator):
    """Generator G(N, seed) of SURVEY.md 8(d), made chemically sane: a bonded
    tree of H/C/N/O/S/F atoms grown with valence caps (so coordination numbers
    stay physical; unphysical CN > ~12 makes the *reference's* autograd return
    NaN through the underflowing Gaussian weights) and a 2.8 Bohr exclusion
    radius towards non-bonded atoms.  Bond lengths U[1.85,2.15] (with H) or
    U[2.45,2.95] Bohr.  Charges: 0.1*N(0,1), shifted to zero sum."""
    z = rng.choice(_ORGANIC_Z, size=nat, p=_ORGANIC_P)
    z[0] = 6
    free = np.array([_VALENCE[int(v)] for v in z])
    xyz = np.zeros((nat, 3))
    for i in range(1, nat):
        cand_parents = np.nonzero(free[:i] > 0)[0]
        if len(cand_parents) == 0:  # saturated: start a new fragment on a carbon
            z[i] = 6
            free[i] = 4
        heavy = cand_parents[z[cand_parents] > 1]
        best, best_d = None, -1.0
        for _ in range(200):
            if len(cand_parents) == 0:
                j, r = int(rng.integers(i)), rng.uniform(3.4, 4.0)
            else:
                j = int(rng.choice(heavy)) if len(heavy) else int(rng.choice(cand_parents))
                r = rng.uniform(1.85, 2.15) if (z[i] == 1 or z[j] == 1) else rng.uniform(2.45, 2.95)
            v = rng.normal(size=3)
            v /= np.linalg.norm(v)
            cand = xyz[j] + r * v
            dist = np.linalg.norm(xyz[:i] - cand, axis=1)
            dist[j] = np.inf
            dmin = dist.min() if i > 1 else np.inf
            if dmin > best_d:
                best, best_d, best_j = cand, dmin, j
            if dmin >= 2.8:
                break
        xyz[i] = best
        if len(cand_parents):
            free[best_j] -= 1
            free[i] -= 1
    q = 0.1 * rng.normal(size=nat)
    q -= q.mean()
    return z.astype(np.int64), xyz, q


def organic_batch(sizes, seed: int):
    """Padded batch of organic blobs: numbers (B,N) int64, positions (B,N,3),
    q (B,N); padding is Z=0, pos=0, q=0."""
    rng = np.random.default_rng(seed)
    nmax = int(max(sizes))
    b = len(sizes)
    numbers = np.zeros((b, nmax), dtype=np.int64)
    pos = np.zeros((b, nmax, 3))
    q = np.zeros((b, nmax))
    for n, nat in enumerate(sizes):
        z, xyz, qq = organic_blob(int(nat), rng)
        numbers[n, :nat], pos[n, :nat], q[n, :nat] = z, xyz, qq
    torch = _torch()
    return torch.from_numpy(numbers), torch.from_numpy(pos), torch.from_numpy(q)


def _blob_job(args):
    nat, seedseq = args
    return organic_blob(int(nat), np.random.default_rng(seedseq))


def organic_batch_parallel(sizes, seed: int, workers: int | None = None):
    """Like :func:`organic_batch` but every structure has its own spawned seed,
    so the batch is reproducible for any worker count (used by bench.py)."""
    import os
    from concurrent.futures import ProcessPoolExecutor

    sizes = [int(s) for s in sizes]
    seeds = np.random.SeedSequence(seed).spawn(len(sizes))
    workers = workers or min(32, os.cpu_count() or 1)
    if workers > 1 and len(sizes) >= 64:
        # fork-based pool: bench.py generates its inputs BEFORE it initialises CUDA / NCCL, so the
        # workers never inherit a device context; the pool is joined when the block exits
        with ProcessPoolExecutor(workers) as ex:
            blobs = list(ex.map(_blob_job, zip(sizes, seeds), chunksize=32))
    else:
        blobs = [_blob_job(a) for a in zip(sizes, seeds)]
    nmax = max(sizes)
    numbers = np.zeros((len(sizes), nmax), dtype=np.int64)
    pos = np.zeros((len(sizes), nmax, 3))
    q = np.zeros((len(sizes), nmax))
    for n, (z, xyz, qq) in enumerate(blobs):
        numbers[n, : len(z)], pos[n, : len(z)], q[n, : len(z)] = z, xyz, qq
    torch = _torch()
    return torch.from_numpy(numbers), torch.from_numpy(pos), torch.from_numpy(q)


# --------------------------------------------------------------------------
# C4: water cluster
# --------------------------------------------------------------------------
def water_cluster(nmol: int, seed: int, rod: tuple[int, int] | None = None):
    """O on a jittered simple-cubic lattice (5.86 Bohr) clipped to a
    sphere, random orientation per molecule, r_OH = 1.81 Bohr, HOH = 104.5 deg.
    ``rod=(a, b)``: an a x b x ceil(nmol / ab) column of the same lattice instead of the sphere
    (every cutoff is exceeded along the column; used by the large-structure parity tests).
    Returns numbers (3 nmol,), positions (3 nmol, 3), q (3 nmol,)."""
    torch = _torch()
    rng = np.random.default_rng(seed)
    if rod is not None:
        a, b = rod
        c = -(-nmol // (a * b))
        grid = np.stack(np.meshgrid(np.arange(c), np.arange(b), np.arange(a), indexing="ij"), -1).reshape(-1, 3)
        grid = grid[:nmol, ::-1].astype(float)
    else:
        m = int(np.ceil((nmol * 6 / np.pi) ** (1 / 3))) + 2
        grid = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3) - (m - 1) / 2
        grid = grid[np.argsort(np.linalg.norm(grid, axis=1), kind="stable")][:nmol]
    o = grid * 5.86 + rng.normal(scale=0.3, size=(nmol, 3))
    a = np.deg2rad(104.5) / 2
    h1 = np.array([np.sin(a), np.cos(a), 0.0]) * 1.81
    h2 = np.array([-np.sin(a), np.cos(a), 0.0]) * 1.81
    # random rotations from normalised quaternions
    qn = rng.normal(size=(nmol, 4))
    qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    w, x, y, z = qn.T
    R = np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1),
    ], 1)  # fmt: skip
    pos = np.stack([o, o + R @ h1, o + R @ h2], 1).reshape(-1, 3)
    numbers = np.tile(np.array([8, 1, 1]), nmol)
    q = np.tile(np.array([-0.66, 0.33, 0.33]), nmol) + 0.02 * rng.normal(size=3 * nmol)
    q -= q.mean()
    return torch.from_numpy(numbers), torch.from_numpy(pos), torch.from_numpy(q)


# --------------------------------------------------------------------------
# C1: the molecule of the reference's examples/single.py:7-27
# --------------------------------------------------------------------------
SINGLE_Z = [6, 6, 6, 6, 7, 6, 16, 1, 1, 1, 1, 1]
SINGLE_XYZ = [
    [-2.56745685564671, -0.02509985979910, 0.0], [-1.39177582455797, +2.27696188880014, 0.0],
    [+1.27784995624894, +2.45107479759386, 0.0], [+2.62801937615793, +0.25927727028120, 0.0],
    [+1.41097033661123, -1.99890996077412, 0.0], [-1.17186102298849, -2.34220576284180, 0.0],
    [-2.39505990368378, -5.22635838332362, 0.0], [+2.41961980455457, -3.62158019253045, 0.0],
    [-2.51744374846065, +3.98181713686746, 0.0], [+2.24269048384775, +4.24389473203647, 0.0],
    [+4.66488984573956, +0.17907568006409, 0.0], [-4.60044244782237, -0.17794734637413, 0.0],
]  # fmt: skip


def single_molecule():
    torch = _torch()
    numbers = torch.tensor([SINGLE_Z])
    positions = torch.tensor([SINGLE_XYZ], dtype=torch.float64)
    q = 0.1 * torch.randn(numbers.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    return numbers, positions, q - q.mean()
