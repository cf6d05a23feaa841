"""Synthetic inputs of the named shapes: self-avoiding random coils with
the algorithm of the reference's generator (`utils/random_xyz.py:41-75`: bond 5.84 A, angle
2.643 rad, uniform random dihedral, 12 A self-avoidance with a 10-bead back-off), chains
placed at uniform random offsets and orientations in the periodic box with a 12 A inter-chain
clash test."""
from __future__ import annotations

import numpy as np

from . import fixtures as fx
from .system import SystemSpec

BL, BA, EV = 5.84, 2.643, 12.0


def random_coil(n: int, rng: np.random.Generator) -> np.ndarray:
    xyz = np.zeros((n, 3))
    xyz[1] = [BL, 0.0, 0.0]
    xyz[2] = xyz[1] + BL * np.array([-np.cos(BA), np.sin(BA), 0.0])
    i, fails = 3, 0
    while i < n:
        b1, b2 = xyz[i - 2] - xyz[i - 3], xyz[i - 1] - xyz[i - 2]
        nrm = np.cross(b1, b2)
        nrm /= np.linalg.norm(nrm)
        u = b2 / np.linalg.norm(b2)
        m = np.cross(nrm, u)
        phi = rng.uniform(-np.pi, np.pi)
        d = -np.cos(BA) * u + np.sin(BA) * (np.cos(phi) * m + np.sin(phi) * nrm)
        cand = xyz[i - 1] + BL * d
        if i > 3 and (np.linalg.norm(xyz[: i - 2] - cand, axis=1) < EV).any():
            fails += 1
            if fails > 50:
                i = max(3, i - 10)   # back off 10 beads
                fails = 0
            continue
        xyz[i] = cand
        i += 1
        fails = 0
    return xyz


def random_rotation(rng: np.random.Generator) -> np.ndarray:
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                     [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                     [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])


def place_chains(nchain: int, nbead: int, box: float, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    cell = EV
    ng = max(1, int(box / cell))
    grid: dict[tuple, list] = {}
    out = np.zeros((nchain * nbead, 3))

    def clashes(pts):
        w = np.mod(pts, box)
        keys = np.floor(w / (box / ng)).astype(int) % ng
        for p, k in zip(w, keys):
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dz in (-1, 0, 1):
                        for q in grid.get(((k[0] + dx) % ng, (k[1] + dy) % ng, (k[2] + dz) % ng), ()):
                            dlt = p - q
                            dlt -= box * np.round(dlt / box)
                            if dlt @ dlt < EV * EV:
                                return True
        return False

    for c in range(nchain):
        while True:
            coil = random_coil(nbead, rng)
            coil = (coil - coil.mean(axis=0)) @ random_rotation(rng).T + rng.uniform(0, box, size=3)
            if not clashes(coil):
                break
        w = np.mod(coil, box)
        keys = np.floor(w / (box / ng)).astype(int) % ng
        for p, k in zip(w, keys):
            grid.setdefault(tuple(k), []).append(p)
        out[c * nbead:(c + 1) * nbead] = coil
    return out


def cag47_box(nchain: int, box: float, seed: int, **kw) -> SystemSpec:
    """C4 (64 chains, 809.972 A, seed 1) / C5 (1024 chains, 2041.0 A, seed 2): (CAG)47 chains,
    rna_cg1.ff, model 1 without MC, DH 0.15 M cutoff 50 A, T 310.15 K, dt 0.5, PBC."""
    spec = fx.cag47(1, pbc=False, **kw)
    spec.seqs = spec.seqs * nchain
    spec.xyz = place_chains(nchain, 141, box, seed)
    spec.box = (box, box, box)
    return spec


def c4() -> SystemSpec:
    return cag47_box(64, 809.972, seed=1)


def c5() -> SystemSpec:
    return cag47_box(1024, 2041.0, seed=2)
