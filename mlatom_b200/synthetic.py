"""Synthetic KREG workloads: fixed equilibrium geometries, Gaussian-displaced
training / test geometries, and labels from an analytic harmonic pair potential so that
energies and gradients are mutually consistent.  numpy only; seeded; no I/O.
"""
from __future__ import annotations

import numpy as np

# Ethanol-shaped 9-atom geometry (C, C, O, H x6), Angstrom.
ETHANOL_EQ = np.array(
    [
        [0.0072, -0.5687, 0.0000],
        [-1.2854, 0.2499, 0.0000],
        [1.1304, 0.3147, 0.0000],
        [0.0392, -1.1972, 0.8900],
        [0.0392, -1.1972, -0.8900],
        [-1.3175, 0.8784, 0.8900],
        [-1.3175, 0.8784, -0.8900],
        [-2.1422, -0.4239, 0.0000],
        [1.9857, -0.1365, 0.0000],
    ],
    dtype=np.float64,
)
ETHANOL_Z = np.array([6, 6, 8, 1, 1, 1, 1, 1, 1], dtype=np.int32)

# Aspirin-shaped 21-atom geometry (C9 H8 O4), Angstrom; any fixed geometry with min pair
# distance >= 0.9 A is acceptable for the workload.
ASPIRIN_EQ = np.array(
    [
        [1.2333, 0.5540, 0.7792],
        [-0.6952, -2.7148, -0.7502],
        [0.7958, -2.1843, 0.8685],
        [1.7813, 0.8105, -1.4821],
        [-0.0857, 0.6088, 0.4403],
        [-0.7927, -0.5515, 0.1244],
        [-0.7288, 1.8464, 0.4133],
        [-2.1426, -0.4741, -0.2184],
        [-2.0787, 1.9238, 0.0706],
        [-2.7855, 0.7636, -0.2453],
        [-0.1409, -1.8536, 0.1477],
        [2.1094, 0.6715, -0.3113],
        [3.5305, 0.5996, 0.1635],
        [-0.1851, 2.7545, 0.6593],
        [-2.7247, -1.3605, -0.4564],
        [-2.5797, 2.8872, 0.0506],
        [-3.8374, 0.8238, -0.5090],
        [3.7290, 1.4184, 0.8593],
        [4.2045, 0.6969, -0.6924],
        [3.7105, -0.3659, 0.6426],
        [-0.2555, -3.5916, -0.7337],
    ],
    dtype=np.float64,
)
ASPIRIN_Z = np.array([8, 8, 8, 8, 6, 6, 6, 6, 6, 6, 6, 6, 6, 1, 1, 1, 1, 1, 1, 1, 1], dtype=np.int32)


# three more hydrogens around the aspirin frame (>= 1.1 Angstrom from every other atom): 22-24 atom test molecules
EXTRA_EQ = np.array([[5.3045, 0.1554, 0.3], [-4.9374, 0.5554, -0.2], [0.0982, 4.0372, 0.5]])


def equilibrium(natoms: int) -> np.ndarray:
    """Equilibrium geometry for `natoms` atoms: ethanol (9), aspirin (21), the first `natoms` atoms of aspirin for
    smaller sizes, aspirin plus up to three far hydrogens for 22-24 (edge-case tests)."""
    if natoms == 9:
        return ETHANOL_EQ.copy()
    if natoms == 21:
        return ASPIRIN_EQ.copy()
    if 2 <= natoms < 21:
        return ASPIRIN_EQ[:natoms].copy()
    if 21 < natoms <= 24:
        return np.vstack([ASPIRIN_EQ, EXTRA_EQ[: natoms - 21]])
    if natoms > 24:
        return lattice_equilibrium(natoms)
    raise ValueError("no synthetic equilibrium geometry for %d atoms" % natoms)


def lattice_equilibrium(natoms: int, spacing: float = 1.5, jitter: float = 0.15) -> np.ndarray:
    """Molecules above 24 atoms: the first `natoms` sites of a cubic lattice (1.5 Angstrom), each displaced by a fixed
    pseudo-random vector of at most `jitter` per component -- every pair distance stays above 1.0 Angstrom."""
    side = int(np.ceil(natoms ** (1.0 / 3.0)))
    grid = np.array([(i, j, k) for i in range(side) for j in range(side) for k in range(side)], dtype=np.float64)
    rng = np.random.default_rng(1000 + natoms)
    return spacing * grid[:natoms] + jitter * (2.0 * rng.random((natoms, 3)) - 1.0)


def atomic_numbers(natoms: int) -> np.ndarray:
    if natoms == 9:
        return ETHANOL_Z.copy()
    if natoms > 24:
        return np.where(np.arange(natoms) % 3 == 0, 6, 1).astype(ASPIRIN_Z.dtype)
    if natoms > 21:
        return np.concatenate([ASPIRIN_Z, np.ones(natoms - 21, dtype=ASPIRIN_Z.dtype)])
    return ASPIRIN_Z[:natoms].copy()


def geometries(eq: np.ndarray, n: int, seed: int, spread: float = 0.05) -> np.ndarray:
    """`n` geometries `eq + N(0, spread)` (Angstrom), shape (n, Nat, 3), float64."""
    rng = np.random.default_rng(seed)
    return eq[None, :, :] + spread * rng.standard_normal((n,) + eq.shape)


def pair_potential(xyz: np.ndarray, eq: np.ndarray, k_seed: int = 7):
    """E = E0 + sum_{a<b} 1/2 k_ab (R_ab - Req_ab)^2 with fixed random k_ab in [0.5, 1.5];
    returns (E (n,), dE/dxyz (n, Nat, 3)).  E0 = -100 keeps |E| away from zero so relative
    energy errors are meaningful."""
    nat = eq.shape[0]
    iu = np.triu_indices(nat, 1)
    kab = 0.5 + np.random.default_rng(k_seed).random(len(iu[0]))
    d_eq = np.linalg.norm(eq[iu[0]] - eq[iu[1]], axis=-1)
    diff = xyz[:, iu[0], :] - xyz[:, iu[1], :]
    r = np.linalg.norm(diff, axis=-1)
    e = -100.0 + 0.5 * np.sum(kab * (r - d_eq) ** 2, axis=1)
    f = (kab * (r - d_eq) / r)[:, :, None] * diff
    g = np.zeros_like(xyz)
    np.add.at(g, (slice(None), iu[0]), f)
    np.add.at(g, (slice(None), iu[1]), -f)
    return e, g
