"""XDM dispersion over the GPU density (density-on-grid consumers): mirror of occ::xdm::XDM
(include/occ/xdm/xdm.h, src/xdm/xdm.cpp, src/xdm/becke_hole.cpp).

What runs where: rho, grad rho, laplacian and tau of the molecular density at every grid point — evaluate_density<2>
(src/gto/density.cpp:72-101), the O(N_grid nbf^2) part — run on the GPU through occ_xc_density_on_grid; the
Becke–Roussel hole, the O(N_grid N_atoms) Hirshfeld-weighted moment sums (xdm.cpp:218-262) and the O(N_atoms^2) pair
energy / forces (xdm.cpp:124-207) stay on the host.  As in occ_b200.hirshfeld the spherical pro-atom densities (the
reference's Thakkar Slater tables, occ::slater) and the free-atom polarizabilities (occ::core::Element) are ARGUMENTS,
so the reference's own tables plug in unchanged.

Reference conventions kept: the density is evaluated from mo.D (restricted: half the electron density) and doubled in
the kernel (fac = 2 rho w, tau = 2 rho[:, 5]); unrestricted adds the alpha and beta blocks first (:264-277); the
analytic Becke-Roussel inversion (becke_hole.cpp:15-68) is the default, hole normalisation 1; points with a
promolecular density below 1e-30 are skipped (:228-230).
"""
from __future__ import annotations

import numpy as np

from . import dft
from .aobasis import AOBasis

ANGSTROM_TO_BOHR = 1.0 / 0.52917721092


def becke_hole_br89_analytic(rho, Q, norm=1.0):
    """becke_hole_br89_analytic (src/xdm/becke_hole.cpp:15-68), vectorised: returns b = x / alpha"""
    rho, Q = np.asarray(rho, dtype=np.float64), np.asarray(Q, dtype=np.float64)
    a1, a2, a3 = 1.5255251812009530, 0.4576575543602858, 0.4292036732051034
    c = [0.7566445420735584, -2.6363977871370960, 5.4745159964232880, -12.657308127108290, 4.1250584725121360,
         -30.42513395716384]
    b = [0.4771976183772063, -1.7799813494556270, 3.8433841862302150, -9.5912050880518490, 2.1730180285916720,
         -30.425133851603660]
    d = [0.00004435009886795587, 0.58128653604457910, 66.742764515940610, 434.26780897229770, 824.7765766052239000,
         1657.9652731582120]
    e = [0.00003347285060926091, 0.47917931023971350, 62.392268338574240, 463.14816427938120, 785.2360350104029000,
         1657.962968223273000000]
    B = 2.085749716493756
    with np.errstate(all="ignore"):
        y = (2.0 / 3.0) * (np.pi * rho / norm) ** (2.0 / 3.0) * rho / Q

        def poly(co, t):
            acc, ti = np.zeros_like(t), np.ones_like(t)
            for k in co:           # ascending powers, accumulated like the reference's loop
                acc = acc + k * ti
                ti = ti * t
            return acc

        neg = y <= 0
        yn = np.where(neg, y, -1.0)
        xn = (-np.arctan(a1 * yn + a2) + a3) * poly(c, yn) / poly(b, yn)
        yp = np.where(neg, 1.0, y)
        By = B * yp
        xp = (np.log(1.0 / By + np.sqrt(1.0 / (By * By) + 1.0)) + 2.0) * poly(d, yp) / poly(e, yp)
        x = np.where(neg, xn, xp)
        alf = (8.0 * np.pi * (rho / np.exp(-x)) / norm) ** (1.0 / 3.0)
        return x / alf


def xdm_dispersion_energy(positions, polarizabilities, moments, a1: float = 0.7, a2: float = 1.4):
    """xdm_dispersion_energy (xdm.cpp:124-207): positions (natoms, 3) Bohr, moments (3, natoms) = <M1^2>, <M2^2>, <M3^2>.
    Returns (energy, forces (natoms, 3)) — Becke-Johnson damped C6 / C8 / C10 sum over atom pairs i < j."""
    pos = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
    pol = np.asarray(polarizabilities, dtype=np.float64)
    m = np.asarray(moments, dtype=np.float64)
    n = len(pol)
    e, forces = 0.0, np.zeros((n, 3))
    for i in range(n):
        for j in range(i + 1, n):
            v = pos[j] - pos[i]
            rij = float(np.linalg.norm(v))
            if rij < 1e-15:
                continue
            factor = pol[i] * pol[j] / (m[0, i] * pol[j] + m[0, j] * pol[i])
            r2 = rij * rij
            r4, r6 = r2 * r2, r2 * r2 * r2
            r8, r10 = r4 * r4, r4 * r6
            c6 = factor * m[0, i] * m[0, j]
            c8 = 1.5 * factor * (m[0, i] * m[1, j] + m[1, i] * m[0, j])
            c10 = 2.0 * factor * (m[0, i] * m[2, j] + m[2, i] * m[0, j]) + 4.2 * factor * m[1, i] * m[1, j]
            rc = (np.sqrt(c8 / c6) + np.sqrt(np.sqrt(c10 / c6)) + np.sqrt(c10 / c8)) / 3.0
            rv = a1 * rc + a2 * ANGSTROM_TO_BOHR
            v2 = rv * rv
            v4, v6 = v2 * v2, v2 * v2 * v2
            v8, v10 = v4 * v4, v4 * v6
            e -= c6 / (v6 + r6) + c8 / (v8 + r8) + c10 / (v10 + r10)
            com = 6.0 * c6 * r4 / (v6 + r6) ** 2 + 8.0 * c8 * r6 / (v8 + r8) ** 2 + 10.0 * c10 * r8 / (v10 + r10) ** 2
            forces[i] += com * v
            forces[j] -= com * v
    return e, forces


class XDM:
    def __init__(self, basis: AOBasis, atom_grids, pro_atom_density, free_polarizabilities, a1: float = 0.7,
                 a2: float = 1.4, device: int = -1):
        """atom_grids: list of (points, weights) per atom (MolecularGrid::get_partitioned_atom_grid, xdm.cpp:283-285);
        pro_atom_density(i, r): spherical free-atom density of atom i; free_polarizabilities (natoms,): Element(n).polarizability"""
        self.basis, self.device, self.a1, self.a2 = basis, device, float(a1), float(a2)
        self.pos = np.asarray(basis.positions, dtype=np.float64).reshape(-1, 3)
        self.Z = np.asarray(basis.atomic_numbers)
        self.free_pol = np.asarray(free_polarizabilities, dtype=np.float64)
        self.pts = np.vstack([np.asarray(p, dtype=np.float64).reshape(-1, 3) for p, _ in atom_grids])
        self.w = np.concatenate([np.asarray(w, dtype=np.float64) for _, w in atom_grids])
        self.r = np.linalg.norm(self.pts[:, None, :] - self.pos[None, :, :], axis=2)
        self.rho_pro = np.stack([np.asarray(pro_atom_density(i, self.r[:, i])) for i in range(len(self.Z))], axis=1)
        self.volume_free = (self.rho_pro * self.w[:, None] * self.r ** 3).sum(0)             # xdm.cpp:421-423

    def populate_moments(self, D: np.ndarray, spinorbital_kind: str = dft.RESTRICTED, ecp_electrons=None):
        """populate_moments_impl + xdm_moment_kernel_restricted / _unrestricted (xdm.cpp:218-277, 372-489); D = mo.D"""
        nspin = 2 if spinorbital_kind == dft.UNRESTRICTED else 1
        rho = dft.evaluate_density_on_grid(self.basis, D, self.pts, 2, spinorbital_kind, device=self.device)
        n = len(self.w)
        rho = rho[:n] + rho[n:] if nspin == 2 else rho
        protot = self.rho_pro.sum(1)
        keep = protot >= 1e-30
        hw = np.zeros_like(self.rho_pro)
        hw[keep] = self.rho_pro[keep] / protot[keep, None]
        r0 = rho[:, 0]
        sigma = (rho[:, 1:4] ** 2).sum(1)
        dsigs = 2.0 * rho[:, 5] - 0.25 * sigma / np.maximum(r0, 1e-30)
        q = (rho[:, 4] - 2.0 * dsigs) / 6.0
        bh = becke_hole_br89_analytic(r0, q, 1.0)
        rsb = np.maximum(self.r - bh[:, None], 0.0)
        wt = np.where(keep, 2.0 * r0 * self.w, 0.0)[:, None] * hw
        self.moments = np.stack([(wt * (self.r ** k - rsb ** k) ** 2).sum(0) for k in (1, 2, 3)])
        self.volume = (wt * self.r ** 3).sum(0)
        ecp = np.zeros(len(self.Z)) if ecp_electrons is None else np.asarray(ecp_electrons, dtype=np.float64)
        self.hirshfeld_charges = self.Z.astype(np.float64) - ecp - wt.sum(0)
        self.num_electrons = float(np.where(keep, 2.0 * r0 * self.w, 0.0).sum())
        self.num_electrons_promolecule = float((protot * self.w)[keep].sum())
        self.polarizabilities = self.volume * self.free_pol / self.volume_free                # xdm_polarizability
        return self.moments

    def energy(self, D: np.ndarray, spinorbital_kind: str = dft.RESTRICTED) -> float:
        """XDM::energy(mo)"""
        self.populate_moments(D, spinorbital_kind)
        self.m_energy, self.m_forces = xdm_dispersion_energy(self.pos, self.polarizabilities, self.moments, self.a1, self.a2)
        return self.m_energy

    def forces(self, D: np.ndarray, spinorbital_kind: str = dft.RESTRICTED) -> np.ndarray:
        """XDM::forces(mo) — (natoms, 3)"""
        self.energy(D, spinorbital_kind)
        return self.m_forces
