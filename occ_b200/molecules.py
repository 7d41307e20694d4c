"""Geometries and synthetic density matrices for the BASELINE.json configs.

water / benzene are analytic; (H2O)32, valinomycin and taxol are seeded, composition-exact synthetic
stand-ins (neither geometry ships with the reference and there is no network) frozen as .xyz files
under occ_b200/data/geometries by scripts/make_geometries.py so every run uses identical coordinates.
The density matrices are integral-free: localised random orbitals, QR-orthonormalised, P = C C^T
(symmetric PSD => rho >= 0), handed to the build as D2 = P (i.e. mo.D = P/2, dft.h:205).
"""
from __future__ import annotations

import os

import numpy as np

from .aobasis import ANGSTROM_TO_BOHR, DATA_DIR, AOBasis, symbol_to_z

GEOM_DIR = os.path.join(DATA_DIR, "geometries")


def read_xyz(path):
    with open(path) as f:
        lines = f.read().strip().splitlines()
    n = int(lines[0])
    Z, pos = [], []
    for ln in lines[2:2 + n]:
        t = ln.split()
        Z.append(symbol_to_z(t[0]))
        pos.append([float(x) for x in t[1:4]])
    return np.array(Z, dtype=np.int32), np.array(pos) * ANGSTROM_TO_BOHR


def water():
    """reference test geometry in Bohr (tests/dft_tests.cpp:1371-1375)"""
    return np.array([8, 1, 1], dtype=np.int32), np.array([[0.0, 0.0, 0.0], [1.43, 1.11, 0.0], [-1.43, 1.11, 0.0]])


def water_isolated_xc():
    """geometry of tests/isolated_xc_tests.cpp:25-29 (Angstrom there)"""
    pos = np.array([[0.0, 0.0, 0.0], [0.0, -0.757, 0.587], [0.0, 0.757, 0.587]]) * ANGSTROM_TO_BOHR
    return np.array([8, 1, 1], dtype=np.int32), pos


def benzene():
    rcc, rch = 1.39, 1.09
    Z, pos = [], []
    for k in range(6):
        a = np.pi / 3 * k
        Z.append(6)
        pos.append([rcc * np.cos(a), rcc * np.sin(a), 0.0])
    for k in range(6):
        a = np.pi / 3 * k
        Z.append(1)
        pos.append([(rcc + rch) * np.cos(a), (rcc + rch) * np.sin(a), 0.0])
    return np.array(Z, dtype=np.int32), np.array(pos) * ANGSTROM_TO_BOHR


def thionyl_chloride_like():
    """SOCl2-like test molecule with second-row atoms (Z > 10: the third alpha set of the NWChem pruning and the larger
    Bragg radii, src/numint/grid_utils.cpp:22-41,234-274), plus a sodium to reach the s/p-only Z = 11 row"""
    Z = np.array([16, 8, 17, 17, 11], dtype=np.int32)
    pos = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 1.44], [1.62, 0.0, -0.95], [-0.81, 1.40, -0.95], [-2.6, -2.3, 0.4]])
    return Z, pos * ANGSTROM_TO_BOHR


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                     [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                     [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])


def generate_water_cluster(nx=4, ny=4, nz=2, spacing=3.1, seed=32):
    """32 rigid waters on a jittered lattice, random orientations (Angstrom)."""
    rng = np.random.default_rng(seed)
    roh, ang = 0.9572, np.deg2rad(104.52)
    mol = np.array([[0.0, 0.0, 0.0], [roh * np.sin(ang / 2), roh * np.cos(ang / 2), 0.0],
                    [-roh * np.sin(ang / 2), roh * np.cos(ang / 2), 0.0]])
    Z, pos = [], []
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                centre = np.array([i, j, k]) * spacing
                for _ in range(1000):
                    c = centre + rng.uniform(-0.25, 0.25, 3)
                    xyz = mol @ _random_rotation(rng).T + c
                    if not pos or np.min(np.linalg.norm(np.array(pos)[None] - xyz[:, None], axis=2)) >= 1.6:
                        break
                else:
                    raise RuntimeError("could not place water")
                pos += list(xyz)
                Z += [8, 1, 1]
    return np.array(Z, dtype=np.int32), np.array(pos)


def generate_globule(n_c, n_h, n_n, n_o, radius, seed):
    """Composition-exact pseudo-molecule: branching self-avoiding walk of heavy atoms (bonds 1.45-1.55 A)
    inside a sphere, hydrogens at 1.09 A, every non-bonded distance >= 1.5 A. Angstrom."""
    rng = np.random.default_rng(seed)
    heavy = [6] * n_c + [7] * n_n + [8] * n_o
    heavy = list(rng.permutation(heavy))
    pos = [np.zeros(3)]

    def unit():
        v = rng.normal(size=3)
        return v / np.linalg.norm(v)

    for _ in heavy[1:]:
        for _try in range(20000):
            parent = int(rng.integers(len(pos)))
            p = pos[parent] + unit() * rng.uniform(1.45, 1.55)
            if np.linalg.norm(p) > radius:
                continue
            d = np.linalg.norm(np.array(pos) - p, axis=1)
            d[parent] = 9.0
            if d.min() >= 1.5:
                pos.append(p)
                break
        else:
            raise RuntimeError("heavy-atom walk failed")
    Z = list(heavy)
    nheavy = len(heavy)
    hcount = [0] * nheavy
    for _ in range(n_h):
        for _try in range(20000):
            parent = int(rng.integers(nheavy))
            if hcount[parent] >= 3:
                continue
            p = pos[parent] + unit() * 1.09
            d = np.linalg.norm(np.array(pos) - p, axis=1)
            d[parent] = 9.0
            if d.min() >= 1.5:
                pos.append(p)
                Z.append(1)
                hcount[parent] += 1
                break
        else:
            raise RuntimeError("hydrogen placement failed")
    return np.array(Z, dtype=np.int32), np.array(pos)


def water_cluster_32():
    return read_xyz(os.path.join(GEOM_DIR, "h2o_32.xyz"))


def valinomycin():
    return read_xyz(os.path.join(GEOM_DIR, "valinomycin_standin.xyz"))


def taxol():
    return read_xyz(os.path.join(GEOM_DIR, "taxol_standin.xyz"))


MOLECULES = {"socl2_na": thionyl_chloride_like, "water": water, "water_isolated": water_isolated_xc, "benzene": benzene, "h2o_32": water_cluster_32,
             "valinomycin": valinomycin, "taxol": taxol}


def synthetic_orbitals(basis: AOBasis, nspin: int = 1, seed: int | None = None, nocc: int | None = None):
    """The localised, orthonormalised random orbitals behind synthetic_density: a list of nspin matrices C_s
    (nbf x nocc, Fortran order) with D2_s = C_s C_s^T (the 1/2 of the unrestricted blocks folded in as 1/sqrt 2)."""
    return [np.asfortranarray(Q * np.sqrt(0.5)) if nspin == 2 else Q for Q in _orthonormal_orbitals(basis, nspin, seed, nocc)]


def _orthonormal_orbitals(basis, nspin, seed, nocc):
    nbf = basis.nbf
    natoms = len(basis.atomic_numbers)
    nocc = basis.num_electrons() // 2 if nocc is None else nocc
    rng = np.random.default_rng(natoms if seed is None else seed)
    centres = basis.positions[basis.bf_to_atom()]
    out = []
    for _ in range(nspin):
        A = rng.normal(size=(nbf, nocc))
        c = rng.integers(natoms, size=nocc)
        d2 = ((centres[:, None, :] - basis.positions[c][None, :, :]) ** 2).sum(-1)
        A *= np.exp(-d2 / 16.0)
        Q, _ = np.linalg.qr(A)
        out.append(np.asfortranarray(Q))
    return out


def synthetic_density(basis: AOBasis, nspin: int = 1, seed: int | None = None, nocc: int | None = None) -> np.ndarray:
    """SURVEY.md §8d recipe. Returns D2 with shape (nspin*nbf, nbf), Fortran (column-major) order,
    un-normalised (callers rescale by N_el / integral(rho), which is linear in D)."""
    blocks = []
    for Q in _orthonormal_orbitals(basis, nspin, seed, nocc):
        P = Q @ Q.T
        P = 0.5 * (P + P.T)
        if nspin == 2:
            P *= 0.5
        blocks.append(P)
    return np.asfortranarray(np.vstack(blocks))
