"""Host-side mirror of occ::gto::Shell / occ::gto::AOBasis (reference: include/occ/gto/shell.h:18-191).

Loads MolSSI-BSE JSON basis sets the way AOBasis::load does (src/gto/shell.cpp:499-571: one Shell per
coefficient column, general/SP contractions split, Cartesian by default with set_pure flipping every
shell) and applies Shell::incorporate_shell_norm (src/gto/shell.cpp:180-186: libcint-style radial
primitive norm, then contracted normalisation; the angular factor is applied at evaluation time).
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

ELEMENT_SYMBOLS = ["X", "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar"]
ANGSTROM_TO_BOHR = 1.0 / 0.52917721067  # include/occ/core/units.h:7-8


def symbol_to_z(sym: str) -> int:
    return ELEMENT_SYMBOLS.index(sym.capitalize())


def _gint(n: int, alpha: float) -> float:
    """src/gto/shell.cpp:20-23"""
    n1_2 = 0.5 * (n + 1)
    return math.gamma(n1_2) / (2.0 * alpha ** n1_2)


def _gto_norm(l: int, alpha: float) -> float:
    """src/gto/shell.cpp:35-50"""
    return 1.0 / math.sqrt(_gint(2 * l + 2, 2.0 * alpha))


def incorporate_shell_norm(l: int, exps: np.ndarray, coefs: np.ndarray) -> np.ndarray:
    """src/gto/shell.cpp:180-186 followed by normalize_contracted_gto (:52-64)."""
    c = np.array([ci * _gto_norm(l, a) for ci, a in zip(coefs, exps)], dtype=np.float64)
    ee = np.array([[_gint(2 * l + 2, ai + aj) for aj in exps] for ai in exps], dtype=np.float64)
    s1 = float(c @ ee @ c)
    return c / math.sqrt(s1)


def shell_size(l: int, pure: bool) -> int:
    return 2 * l + 1 if pure else (l + 1) * (l + 2) // 2


@dataclass
class Shell:
    l: int
    origin: np.ndarray          # Bohr
    exponents: np.ndarray
    coefficients: np.ndarray    # after incorporate_shell_norm
    atom: int


@dataclass
class AOBasis:
    atomic_numbers: np.ndarray   # (natoms,)
    positions: np.ndarray        # (natoms, 3) Bohr
    shells: list = field(default_factory=list)
    pure: bool = False
    name: str = ""

    @staticmethod
    def load(atomic_numbers, positions_bohr, name: str, pure: bool = False) -> "AOBasis":
        path = name if name.endswith(".json") else os.path.join(DATA_DIR, "basis", name.lower() + ".json")
        with open(path) as f:
            elements = json.load(f)["elements"]
        Z = np.asarray(atomic_numbers, dtype=np.int32)
        pos = np.asarray(positions_bohr, dtype=np.float64).reshape(-1, 3)
        basis = AOBasis(Z, pos, [], pure, name)
        for a, z in enumerate(Z):
            if str(int(z)) not in elements:
                raise RuntimeError(f"element {z} not found in basis")  # shell.cpp:524-526
            for s in elements[str(int(z))]["electron_shells"]:
                exps = np.array([float(x) for x in s["exponents"]])
                am = s["angular_momentum"]
                for i, col in enumerate(s["coefficients"]):   # general contractions split (shell.cpp:532-538)
                    l = int(am[i % len(am)])
                    c = np.array([float(x) for x in col])
                    basis.shells.append(Shell(l, pos[a].copy(), exps.copy(), incorporate_shell_norm(l, exps, c), a))
        return basis

    @property
    def nsh(self) -> int:
        return len(self.shells)

    @property
    def first_bf(self) -> np.ndarray:
        out, n = [], 0
        for s in self.shells:
            out.append(n)
            n += shell_size(s.l, self.pure)
        return np.array(out, dtype=np.int32)

    @property
    def nbf(self) -> int:
        return int(sum(shell_size(s.l, self.pure) for s in self.shells))

    @property
    def l_max(self) -> int:
        return max(s.l for s in self.shells)

    def bf_to_atom(self) -> np.ndarray:
        out = []
        for s in self.shells:
            out += [s.atom] * shell_size(s.l, self.pure)
        return np.array(out, dtype=np.int32)

    def num_electrons(self) -> int:
        return int(self.atomic_numbers.sum())

    def as_arrays(self) -> dict:
        """Flat arrays in the layout of occ_xc_basis_t (include/occ_xc_b200.h)."""
        nprim = np.array([len(s.exponents) for s in self.shells], dtype=np.int32)
        off = np.zeros(self.nsh, dtype=np.int32)
        off[1:] = np.cumsum(nprim)[:-1]
        return {
            "nsh": self.nsh,
            "pure": int(self.pure),
            "l": np.array([s.l for s in self.shells], dtype=np.int32),
            "origin": np.ascontiguousarray(np.array([s.origin for s in self.shells], dtype=np.float64)),
            "nprim": nprim,
            "prim_off": off,
            "exps": np.concatenate([s.exponents for s in self.shells]).astype(np.float64),
            "coefs": np.concatenate([s.coefficients for s in self.shells]).astype(np.float64),
            "shell_atom": np.array([s.atom for s in self.shells], dtype=np.int32),
        }
