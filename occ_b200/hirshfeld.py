"""Density-on-grid consumers over the GPU density: the Hirshfeld partition and a cube-grid
density, mirrors of occ::dft::HirshfeldPartition::compute_hirshfeld_weights (src/dft/hirshfeld.cpp:247-476, charges and
atomic volumes) and of the cube writers' evaluate_density_on_grid calls (include/occ/gto/density.h:67-107).

What runs where: the molecular density at every grid point — the O(N_grid nbf^2) part, src/gto/density.cpp through
occ_xc_density_on_grid — runs on the GPU; the pro-atomic densities and the O(N_grid N_atoms) weighted sums stay on the
host.  The reference takes its spherical pro-atom densities from the Thakkar Slater-type tables of occ::slater
(src/slater/thakkar_basis.cpp, an embedded 12 000-line data table outside this path): here they are an argument,
`pro_atom_density(atom_index, r) -> rho(r)`, so the reference's own tables plug in unchanged.

Reference conventions kept:
  * restricted: the kernel uses 2 * rho with rho evaluated from mo.D (hirshfeld.cpp:30,78) — evaluate from D = mo.D;
    unrestricted: rho_alpha + rho_beta of the stacked mo.D, again doubled (:53-58);
  * every atom grid is processed in blocks of 64 points with  u = min(npt - 1, (block + 1) * 64)  (:287-289), i.e. the LAST
    point of each atom grid never enters the sums; `drop_last_point=True` reproduces that, False integrates them all;
  * grid points with a promolecular density below 1e-30 are skipped (:24-26).
"""
from __future__ import annotations

import numpy as np

from . import dft
from .aobasis import AOBasis


class HirshfeldPartition:
    def __init__(self, basis: AOBasis, atom_grids, pro_atom_density, ecp_electrons=None, device: int = -1,
                 drop_last_point: bool = True):
        """atom_grids: list of (points (n_a, 3), weights (n_a,)) per atom — MolecularGrid::get_partitioned_atom_grid(i)
        (hirshfeld.cpp:198-202); pro_atom_density(i, r): spherical free-atom density of atom i at distances r."""
        self.basis, self.device = basis, device
        self.pos = np.asarray(basis.positions, dtype=np.float64).reshape(-1, 3)
        self.Z = np.asarray(basis.atomic_numbers)
        self.ecp = np.zeros(len(self.Z)) if ecp_electrons is None else np.asarray(ecp_electrons, dtype=np.float64)
        pts, w = [], []
        for p, ww in atom_grids:
            p, ww = np.asarray(p, dtype=np.float64).reshape(-1, 3), np.asarray(ww, dtype=np.float64)
            if drop_last_point and len(ww) > 0:
                p, ww = p[:-1], ww[:-1]
            pts.append(p)
            w.append(ww)
        self.pts, self.w = np.vstack(pts), np.concatenate(w)
        self.r = np.linalg.norm(self.pts[:, None, :] - self.pos[None, :, :], axis=2)          # (npts, natoms)
        self.rho_pro = np.stack([np.asarray(pro_atom_density(i, self.r[:, i])) for i in range(len(self.Z))], axis=1)
        # free-atom volumes: integral of rho_pro_i r_i^3 over the whole molecular grid (hirshfeld.cpp:346-349)
        self.free_atom_volumes = (self.rho_pro * self.w[:, None] * self.r ** 3).sum(0)

    def calculate(self, D: np.ndarray, spinorbital_kind: str = dft.RESTRICTED):
        """HirshfeldPartition::calculate(mo): D = mo.D.  Returns the charges; atom_volumes, num_electrons,
        num_electrons_promolecule are left on the object like the reference's members."""
        nspin = 2 if spinorbital_kind == dft.UNRESTRICTED else 1
        rho = dft.evaluate_density_on_grid(self.basis, D, self.pts, 0, spinorbital_kind, device=self.device)[:, 0]
        n = len(self.w)
        rho_tot = rho[:n] + rho[n:] if nspin == 2 else rho
        protot = self.rho_pro.sum(1)
        keep = protot >= 1e-30
        hw = np.zeros_like(self.rho_pro)
        hw[keep] = self.rho_pro[keep] / protot[keep, None]
        fac = np.where(keep, 2.0 * rho_tot * self.w, 0.0)
        self.hirshfeld_charges = self.Z.astype(np.float64) - self.ecp - (hw * fac[:, None]).sum(0)
        self.atom_volumes = (hw * self.r ** 3 * fac[:, None]).sum(0)
        self.num_electrons = float(fac.sum())
        self.num_electrons_promolecule = float((protot * self.w)[keep].sum())
        return self.hirshfeld_charges


def cube_density(basis: AOBasis, D: np.ndarray, origin, steps, axes=None, spinorbital_kind: str = dft.RESTRICTED,
                 device: int = -1) -> np.ndarray:
    """Electron density on a regular (cube-file) grid: origin (3,), steps (nx, ny, nz), axes 3 x 3 step vectors in Bohr
    (rows; default 0.2 Bohr along x, y, z).  D as evaluate_density receives it (callers pass 2 * mo.D).  Returns an array
    of shape (nx, ny, nz) [restricted] or (2, nx, ny, nz) [alpha, beta], x slowest like the cube format."""
    nx, ny, nz = (int(s) for s in steps)
    axes = np.eye(3) * 0.2 if axes is None else np.asarray(axes, dtype=np.float64).reshape(3, 3)
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    pts = np.asarray(origin, dtype=np.float64)[None, :] + i.reshape(-1, 1) * axes[0] + j.reshape(-1, 1) * axes[1] + \
        k.reshape(-1, 1) * axes[2]
    rho = dft.evaluate_density_on_grid(basis, D, pts, 0, spinorbital_kind, device=device)[:, 0]
    nspin = 2 if spinorbital_kind == dft.UNRESTRICTED else 1
    return rho.reshape(nx, ny, nz) if nspin == 1 else rho.reshape(2, nx, ny, nz)
